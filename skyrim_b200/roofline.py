"""Algorithmic work of one 6-h step, per kernel family (the figures bench.py's ``roofline``
object and DESIGN.md use; derivation in SURVEY.md §8(d)).  FLOPs are 2*MAC of the dense
contractions; bytes are the mandatory HBM traffic of the family's fused kernel."""
from __future__ import annotations

from .config import PanguConfig, SFNOConfig

WIN_TOK = 144


def pangu_flops(cfg: PanguConfig) -> dict:
    C = cfg.dim
    HW = cfg.H * cfg.W
    T1, T2 = cfg.Z * HW, cfg.Z * cfg.H2 * cfg.W2
    nwin1 = (cfg.Z // 2) * (cfg.padded_h(cfg.H) // 6) * (cfg.W // 12)
    nwin2 = (cfg.Z // 2) * (cfg.padded_h(cfg.H2) // 6) * (cfg.W2 // 12)
    f = dict(qkv=0.0, attn=0.0, proj=0.0, fc1=0.0, fc2=0.0)
    for li, depth in enumerate(cfg.depths):
        c, T, rows = (C, T1, nwin1 * WIN_TOK) if li in (0, 3) else (2 * C, T2, nwin2 * WIN_TOK)
        f["qkv"] += depth * 2.0 * rows * c * 3 * c
        f["attn"] += depth * 2.0 * 2.0 * rows * WIN_TOK * c
        f["proj"] += depth * 2.0 * rows * c * c
        f["fc1"] += depth * 2.0 * T * c * 4 * c
        f["fc2"] += depth * 2.0 * T * 4 * c * c
    f["mlp"] = f["fc1"] + f["fc2"]
    nzt = cfg.Z - 1
    f["embed"] = 2.0 * nzt * HW * 160 * C + 2.0 * HW * 112 * C
    f["recover"] = 2.0 * nzt * HW * 2 * C * 160 + 2.0 * HW * 2 * C * 64
    f["down"] = 2.0 * T2 * 4 * C * 2 * C
    f["up"] = 2.0 * T2 * 2 * C * 4 * C + 2.0 * T1 * C * C
    f["total"] = (f["qkv"] + f["attn"] + f["proj"] + f["fc1"] + f["fc2"] + f["embed"] + f["recover"]
                  + f["down"] + f["up"])
    return f


def pangu_bytes(cfg: PanguConfig) -> dict:
    """Mandatory HBM bytes per 6-h step and kernel family for the data flow of DESIGN.md section 3 (fp32 residual stream x,
    fp16 operand images xh; weights stay in L2 and are not counted).  Per token and feature:
      qkv    read xh 2 B, write q|k|v window images 3 x 2 B (padded tokens included)
      attn   read q|k|v 6 B, write the projection's operand image 2 B
      proj   read operand image 2 B, residual x read 4 B + write 4 B, next operand image 2 B
      mlp    read xh 2 B, x read + write 8 B, xh write 2 B (the 4C hidden activation never leaves the SM)"""
    C = cfg.dim
    HW = cfg.H * cfg.W
    T1, T2 = cfg.Z * HW, cfg.Z * cfg.H2 * cfg.W2
    nwin1 = (cfg.Z // 2) * (cfg.padded_h(cfg.H) // 6) * (cfg.W // 12)
    nwin2 = (cfg.Z // 2) * (cfg.padded_h(cfg.H2) // 6) * (cfg.W2 // 12)
    b = dict(qkv=0.0, attn=0.0, proj=0.0, mlp=0.0)
    for li, depth in enumerate(cfg.depths):
        c, T, rows = (C, T1, nwin1 * WIN_TOK) if li in (0, 3) else (2 * C, T2, nwin2 * WIN_TOK)
        b["qkv"] += depth * (2.0 * T * c + 6.0 * rows * c)
        b["attn"] += depth * (6.0 * rows * c + 2.0 * T * c)
        b["proj"] += depth * 12.0 * T * c
        b["mlp"] += depth * 12.0 * T * c
    state = float(pangu_state_bytes(cfg))
    b["embed"] = state + T1 * C * (4.0 + 2.0)                  # state in; x (fp32) + operand image out
    b["recover"] = 2.0 * T1 * C * 2.0 + state                  # concat(skip, x) images in; state out
    b["down"] = T1 * C * 4.0 + T2 * 4 * C * 2.0 * 2 + T2 * 2 * C * (4.0 + 2.0)   # x in, merged image out + in, x2 + image out
    b["up"] = T2 * 2 * C * 2.0 + T1 * C * 2.0 * 2 + T1 * C * (4.0 + 2.0)       # image in, shuffled image out + in, x1 + image out
    b["total"] = sum(b.values())
    return b


def pangu_state_bytes(cfg: PanguConfig) -> int:
    return cfg.n_channels * cfg.nlat * cfg.nlon * 4


def sfno_flops(cfg: SFNOConfig) -> dict:
    """SURVEY.md §8(d) formulae (Legendre contractions, per-l channel mixing, pixel MLPs); the
    longitude DFT is counted as an FFT (negligible) although the engine runs it as a GEMM, and
    the 3-term fp16 split triples the executed tensor work without changing these figures."""
    E, Cin, L = cfg.embed, cfg.n_channels, cfg.layers
    P1, P2 = cfg.nlat * cfg.nlon, cfg.h * cfg.w
    lm = cfg.lmax * cfg.mmax
    f = {}
    f["sfno_enc"] = 2.0 * P1 * (Cin * E + E * E)
    f["sfno_dec"] = 2.0 * P1 * ((E + Cin) * E + E * Cin)
    leg_big, leg_int = 4.0 * cfg.nlat * lm * E, 4.0 * cfg.h * lm * E
    f["sfno_sht"] = leg_big + (L - 1) * leg_int
    # inverse transforms: mixed spectrum of every block + the resampled residual of the two grid-changing blocks
    f["sfno_isht"] = (L - 1) * leg_int + leg_big + leg_int + leg_big
    f["sfno_spec"] = L * 8.0 * lm * E * E
    mlp_int = 2.0 * P2 * (E * E + 2 * cfg.mlp_ratio * E * E)
    mlp_big = 2.0 * P1 * (E * E + 2 * cfg.mlp_ratio * E * E)
    f["sfno_mlp"] = (L - 1) * mlp_int + mlp_big
    f["total"] = sum(f.values())
    return f


def sfno_bytes(cfg: SFNOConfig) -> dict:
    """Mandatory HBM bytes per step and family with fp32 activations between stages (E channels; pixels P1 at 721x1440,
    P2 on the internal grid); spectral weights are streamed once per step and layer."""
    E, Cin, L = cfg.embed, cfg.n_channels, cfg.layers
    P1, P2 = cfg.nlat * cfg.nlon, cfg.h * cfg.w
    lm = cfg.lmax * cfg.mmax
    b = {}
    b["sfno_enc"] = 4.0 * P1 * (Cin + E) + 4.0 * P1 * E          # state in, x out, positional embedding in
    b["sfno_dec"] = 4.0 * P1 * (E + Cin + Cin)
    b["sfno_sht"] = 4.0 * E * (P1 + (L - 1) * P2) + 8.0 * L * E * lm
    b["sfno_isht"] = 8.0 * (L + 2) * E * lm + 4.0 * E * (2 * P1 + L * P2)
    b["sfno_spec"] = L * (8.0 * cfg.lmax * E * E + 16.0 * E * lm)
    b["sfno_mlp"] = 4.0 * E * 4 * ((L - 1) * P2 + P1)
    b["total"] = sum(b.values())
    return b
