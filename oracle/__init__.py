"""CPU oracle for the debvader hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference's hot path
(batched conv-VAE deblending of postage stamps + stamp extraction / window
subtract-back).  It is the checker the CUDA path is compared with; it is never
the thing measured or shipped.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
The product package ``debvader_b200`` must not import anything from here.

Parity status
-------------
* extraction / window arithmetic / MSE (integer + copy + fp64 work):
  PINNED.  ``tests/golden/make_golden.py`` runs the reference's own
  ``extract/extraction.py`` and ``deblend/field_deblender.py`` (imported from
  /root/reference in the build container) and commits their outputs under
  ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this oracle against
  them bit-for-bit (extraction) / to 1e-12 (spline-shift residual).
* network (encoder / latent / decoder): PARITY UNPINNED.  The arithmetic lives
  in TensorFlow 2.13.0 / tensorflow-probability 0.21.0 / Keras
  (requirements.txt:9-10 of the reference), none of which is importable in the
  build container, and the reference ships no test or golden vector for network
  outputs.  The restatement follows model/model.py line by line (cited in each
  function) and the published semantics of those libraries; what *is* pinned is
  the architecture (64 tensor shapes from the shipped checkpoint index, the
  parameter totals 3 741 224 / 4 577 228 of the reference's ``net.summary()``)
  and ``fill_triangular`` (the reference's own ONNX twin, model/model.py:43-58).
  Two independent implementations (explicit numpy loops/einsum in
  ``vae_numpy`` and torch-CPU library convolutions in ``vae_torch``) are
  cross-checked against each other.
"""
