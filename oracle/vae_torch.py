"""torch-CPU restatement of the reference network (oracle; test infrastructure).

Same algorithm as ``oracle.vae_numpy`` (see the citations there — reference
model/model.py:43-161) but with the convolutions delegated to torch's CPU
library kernels, so that thousands of stamps finish in seconds and the
``cpu_baseline`` leg of bench.py has an all-cores CPU implementation to time.
It is an independent second implementation: tests cross-check it against the
explicit numpy one.

``emulate`` = None | "bf16" | "fp16": round the weights and every activation
that the GPU tensor-core path stores in 16 bits to that type (fp32 accumulate),
which predicts the error of that path and gives a tight comparison target.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .vae_numpy import BN_EPS, DIAG_SHIFT, SCALE_SHIFT, D, E, same_pad


def _q(t, emulate):
    if emulate is None:
        return t
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[emulate]
    return t.to(dt).to(t.dtype)


class TorchOracle:
    def __init__(self, weights: dict, dtype=torch.float32, emulate=None, threads=None):
        if threads:
            torch.set_num_threads(threads)
        self.dtype = dtype
        self.emulate = emulate
        self.keep_acts = False  # when True, forward() records every layer's NHWC output in self.acts
        self.acts = {}
        self.w = {k: torch.from_numpy(np.asarray(v)).to(dtype) for k, v in weights.items()}
        w = self.w
        self.enc_convs = []
        n = 1
        while (E % (n, "kernel")) in w and w[E % (n, "kernel")].ndim == 4:
            for j, s in ((n, 1), (n + 2, 2)):
                k = _q(w[E % (j, "kernel")], emulate).permute(3, 2, 0, 1).contiguous()  # HWIO -> OIHW
                a = w[E % (j + 1, "alpha")].permute(2, 0, 1).contiguous()  # HWC -> CHW
                self.enc_convs.append((k, w[E % (j, "bias")], a, s))
            n += 4
        self.enc_flat_alpha = w[E % (n, "alpha")]
        self.enc_dense = (_q(w[E % (n + 1, "kernel")], emulate), w[E % (n + 1, "bias")])
        self.dec_convs = []
        n = 5
        while (D % (n + 1, "alpha")) in w:
            for j, s in ((n, 2), (n + 2, 1)):
                # TF (kh,kw,out,in) -> torch conv_transpose2d weight (in,out,kh,kw)
                k = _q(w[D % (j, "kernel")], emulate).permute(3, 2, 0, 1).contiguous()
                a = w[D % (j + 1, "alpha")].permute(2, 0, 1).contiguous()
                self.dec_convs.append((k, w[D % (j, "bias")], a, s))
            n += 4
        self.head = (_q(w[D % (n, "kernel")], emulate).permute(3, 2, 0, 1).contiguous(), w[D % (n, "bias")])

    # ---- encoder: model.py:61-100 -------------------------------------------------
    def encode(self, x: torch.Tensor) -> torch.Tensor:
        w, em = self.w, self.emulate
        x = x.to(self.dtype)
        g, b = w[E % (0, "gamma")], w[E % (0, "beta")]
        m, v = w[E % (0, "moving_mean")], w[E % (0, "moving_variance")]
        h = g * (x - m) / torch.sqrt(v + BN_EPS) + b
        h = _q(h, em).permute(0, 3, 1, 2)  # NCHW (the BatchNorm output is what conv1's im2col stores in 16 bits)
        last = len(self.enc_convs) - 1
        for i, (k, bias, alpha, s) in enumerate(self.enc_convs):
            _, pt, pb = same_pad(h.shape[2], k.shape[2], s)
            _, pl, pr = same_pad(h.shape[3], k.shape[3], s)
            h = F.conv2d(F.pad(h, (pl, pr, pt, pb)), k, bias, stride=s)
            h = torch.clamp_min(h, 0) + alpha * torch.clamp_max(h, 0)
            if i != last:  # the last conv's epilogue also applies the Flatten PReLU before storing
                h = _q(h, em)
                if self.keep_acts:
                    self.acts["enc_conv%d" % (i + 1)] = h.permute(0, 2, 3, 1).contiguous()
        h = h.permute(0, 2, 3, 1).reshape(h.shape[0], -1)  # Flatten (h,w,c)
        h = torch.clamp_min(h, 0) + self.enc_flat_alpha * torch.clamp_max(h, 0)
        h = _q(h, em)
        if self.keep_acts:
            self.acts["enc_conv8"] = h.reshape(-1, 4, 4, 256)
        return h @ self.enc_dense[0] + self.enc_dense[1]

    # ---- latent: model.py:43-58 ---------------------------------------------------
    def latent(self, params: torch.Tensor, eps=None, latent_dim=32):
        loc = params[:, :latent_dim]
        x = params[:, latent_dim:]
        xc = torch.cat([x[:, latent_dim:], torch.flip(x, dims=[1])], dim=1)
        tril = torch.tril(xc.reshape(-1, latent_dim, latent_dim))
        idx = torch.arange(latent_dim)
        tril[:, idx, idx] = F.softplus(tril[:, idx, idx]) + DIAG_SHIFT
        z = loc.clone() if eps is None else loc + torch.einsum("bij,bj->bi", tril, eps.to(self.dtype))
        return {"loc": loc, "scale_tril": tril, "z": z, "stddev": torch.sqrt((tril * tril).sum(-1))}

    # ---- decoder: model.py:103-161 ------------------------------------------------
    def decode(self, z: torch.Tensor):
        w, em = self.w, self.emulate
        pr = lambda t, a: torch.clamp_min(t, 0) + a * torch.clamp_max(t, 0)
        h = pr(z.to(self.dtype), w[D % (0, "alpha")])
        # the GPU tensor-core path keeps Dense(32->560) in fp32 SIMT, so no rounding here
        h = pr(h @ w[D % (1, "kernel")] + w[D % (1, "bias")], w[D % (2, "alpha")])
        h = _q(h, em)
        if self.keep_acts:
            self.acts["dec_dense1"] = h.reshape(-1, 1, 1, h.shape[1])
        h = pr(h @ _q(w[D % (3, "kernel")], em) + w[D % (3, "bias")], w[D % (4, "alpha")])
        h = _q(h, em)
        if self.keep_acts:
            self.acts["dec_dense2"] = h.reshape(-1, 4, 4, 256)
        cin = self.dec_convs[0][0].shape[0]
        ww = int(round((h.shape[1] // cin) ** 0.5))
        h = h.reshape(-1, ww, ww, cin).permute(0, 3, 1, 2)
        for i, (k, bias, alpha, s) in enumerate(self.dec_convs):
            n = h.shape[2]
            if s == 2:
                h = F.conv_transpose2d(h, k, bias, stride=2, padding=0)[:, :, : 2 * n, : 2 * n]
            else:
                h = F.conv_transpose2d(h, k, bias, stride=1, padding=1)
            h = _q(pr(h, alpha), em)
            if self.keep_acts:
                self.acts["dec_convT%d" % (i + 1)] = h.permute(0, 2, 3, 1).contiguous()
        k, bias = self.head
        h = torch.relu(F.conv2d(F.pad(h, (1, 1, 1, 1)), k, bias))
        S = 59
        crop = h.shape[2] - S
        lo = crop // 2
        h = h[:, :, lo : lo + S, lo : lo + S].permute(0, 2, 3, 1)
        C = h.shape[-1] // 2
        return h[..., :C].contiguous(), (SCALE_SHIFT + h[..., C:]).contiguous()

    def forward(self, x, eps=None):
        with torch.no_grad():
            x = torch.as_tensor(x)
            params = self.encode(x)
            lat = self.latent(params, None if eps is None else torch.as_tensor(eps))
            mean, std = self.decode(lat["z"])
        return {"params": params, "z": lat["z"], "z_loc": lat["loc"], "z_stddev": lat["stddev"], "mean": mean, "stddev": std}
