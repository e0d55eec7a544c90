"""debvader_b200 — B200-native implementation of debvader's data-parallel hot path.

Public surface mirrors the reference package (src/debvader/__init__.py:1-2):
``DeblendField`` and ``IterativeDeblendField``; the rest is imported by module path
(``debvader_b200.model.model.load_deblender`` ...).  The ``debvader`` shim package at the repo root
re-exports everything under the reference's own module paths.  Imports are lazy so that the pure
host-side modules (index planning, checkpoint reader, spec) work without a GPU; any compute call
without the CUDA library / a B200 raises.
"""
__version__ = "0.1.0"


def __getattr__(name):
    if name == "DeblendField":
        from .deblend.field_deblender import DeblendField

        return DeblendField
    if name == "IterativeDeblendField":
        from .deblend_iterative.iterative_deblender import IterativeDeblendField

        return IterativeDeblendField
    raise AttributeError(name)
