"""ORACLE — test infrastructure, not product code.

CPU fp32 / fp64 restatement of the FourCastNet-v2-small SFNO 6-hour forward that the
reference reaches through ``fcnv2_sm.load(registry.get_model("e2mip://fcnv2_sm"))``
(/root/reference/skyrim/core/models/fourcastnet_v2.py:36-37) and steps through
earth2mip's ``Inference`` wrapper (normalise, yield IC, x = model(x), de-normalise;
/root/reference/skyrim/core/models/utils.py:34-40).

PARITY UNPINNED: earth2mip (unpinned git HEAD, requirements.txt:2), modulus / makani's
sfnonet and torch-harmonics are not vendored under /root/reference and cannot be imported
here; the reference's tests hold no golden vector for this path.  The network below follows
SURVEY.md Appendix B with the hyper-parameters marked "default synthetic config" there
(E=384, L=8, scale_factor=3) and the free choices fixed in DESIGN.md (inner skip = 1x1 conv,
outer skip = identity, instance norm with affine, no spectral bias).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from skyrim_b200.config import SFNOConfig
from oracle import sht_ref


class SFNORef:
    def __init__(self, cfg: SFNOConfig, weights, dtype=torch.float32, emulate: str | None = None):
        self.cfg, self.dtype, self.emulate = cfg, dtype, emulate
        self.w = {k: torch.from_numpy(np.asarray(v)).to(dtype) for k, v in weights.items()}
        cd = torch.complex128 if dtype == torch.float64 else torch.complex64
        # Legendre / quadrature tables from the oracle's OWN builder (scipy), not the product's skyrim_b200/sht.py
        self.tabs = {}
        for tag, (nlat, nlon, grid) in (("big", (cfg.nlat, cfg.nlon, "equiangular")), ("int", (cfg.h, cfg.w, "legendre-gauss"))):
            fwd, inv, _, _ = sht_ref.tables(nlat, cfg.lmax, cfg.mmax, grid)
            self.tabs[tag] = dict(fwd=torch.from_numpy(fwd).to(dtype), inv=torch.from_numpy(inv).to(dtype), nlon=nlon, nlat=nlat)
        self.cd = cd

    def _q(self, t, tag="all"):
        """operand rounding emulation: emulate = "fp16" rounds every GEMM operand to half; a set / tuple of tags
        ("mlp", "dft", "leg", "spec") rounds only those contractions (precision-budget studies, tests/test_sfno_cpu.py)"""
        if self.emulate == "fp16" or (isinstance(self.emulate, (set, frozenset, tuple, list)) and tag in self.emulate):
            if t.is_complex():
                return torch.complex(t.real.half().to(self.dtype), t.imag.half().to(self.dtype))
            return t.half().to(self.dtype)
        return t

    # -- transforms (channel-major (C, nlat, nlon) <-> complex (C, lmax, mmax)) -----------------
    def sht(self, x, tag):
        t = self.tabs[tag]
        f = 2.0 * np.pi * torch.fft.rfft(self._q(x, "dft"), dim=-1, norm="forward")[..., : self.cfg.mmax]
        f = self._q(f, "leg")
        return torch.einsum("mlk,ckm->clm", self._q(t["fwd"], "leg").to(self.cd), f)

    def isht(self, X, tag):
        t = self.tabs[tag]
        f = torch.einsum("mkl,clm->ckm", self._q(t["inv"], "leg").to(self.cd), self._q(X, "leg"))
        return torch.fft.irfft(self._q(f, "dft"), n=t["nlon"], dim=-1, norm="forward")

    def _conv1x1(self, x, w, b=None):  # x (C, H, W)
        y = torch.einsum("oc,chw->ohw", self._q(w, "mlp"), self._q(x, "mlp"))
        return y if b is None else y + b[:, None, None]

    def _inorm(self, x, g, b):
        return F.instance_norm(x[None], weight=g, bias=b, eps=self.cfg.eps)[0]

    def block(self, x, i):
        cfg, w = self.cfg, self.w
        p = f"blk{i}."
        tin = "big" if i == 0 else "int"
        tout = "big" if i == cfg.layers - 1 else "int"
        xn = self._inorm(x, w[p + "norm0.g"], w[p + "norm0.b"])
        X = self.sht(xn, tin)
        residual = xn if tin == tout else self.isht(X, tout)
        wc = torch.complex(w[p + "spec.w"][..., 0], w[p + "spec.w"][..., 1])  # (l, out, in)
        Xo = torch.einsum("loi,ilm->olm", self._q(wc, "spec"), self._q(X, "spec"))
        y = self.isht(Xo, tout)
        y = y + self._conv1x1(residual, w[p + "inner.w"], w[p + "inner.b"])
        y = F.gelu(y)
        y = self._inorm(y, w[p + "norm1.g"], w[p + "norm1.b"])
        hdn = F.gelu(self._conv1x1(y, w[p + "fc1.w"], w[p + "fc1.b"]))
        y = self._conv1x1(hdn, w[p + "fc2.w"], w[p + "fc2.b"])
        return y + residual

    @torch.no_grad()
    def step(self, state):
        """(73, nlat, nlon) de-normalised -> (73, nlat, nlon) de-normalised."""
        w = self.w
        x = torch.as_tensor(state).to(self.dtype)
        xh = (x - w["norm.mean"][:, None, None]) / w["norm.std"][:, None, None]
        y = F.gelu(self._conv1x1(xh, w["enc.fc1.w"], w["enc.fc1.b"]))
        y = self._conv1x1(y, w["enc.fc2.w"], w["enc.fc2.b"]) + w["pos_embed"]
        for i in range(self.cfg.layers):
            y = self.block(y, i)
        y = torch.cat([y, xh], 0)
        y = F.gelu(self._conv1x1(y, w["dec.fc1.w"], w["dec.fc1.b"]))
        y = self._conv1x1(y, w["dec.fc2.w"], w["dec.fc2.b"])
        return y * w["norm.std"][:, None, None] + w["norm.mean"][:, None, None]
