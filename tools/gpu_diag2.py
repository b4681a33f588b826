"""GPU diagnostics (development): Pangu full-size errors per channel; SFNO parity over a list of shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from skyrim_b200.engine import StepEngine
from skyrim_b200.config import *
from skyrim_b200.weights import *
from skyrim_b200.verify import compare_fullsize, load_fixture, summarise
from oracle.pangu_ref import rel_err_per_channel

what = sys.argv[1:] or ["pangu", "sfno"]
if "pangu" in what:
    fx = load_fixture("pangu")
    cfg = pangu_full()
    x0 = synthetic_state(PANGU_CHANNELS, cfg.nlat, cfg.nlon, 0)
    eng = StepEngine(cfg, 0); eng.load_weights(make_pangu_weights(cfg, 0))
    y = eng.step(torch.from_numpy(x0)[None].cuda())[0]
    c = compare_fullsize(y, fx)
    print("pangu full:", summarise(c))
    for k in ("rel", "nrm", "block", "last"):
        o = np.argsort(-c[k])[:6]
        print(" ", k, [(PANGU_CHANNELS[i], float(f"{c[k][i]:.3e}")) for i in o])
    # where are the block errors? (lat-block, lon-block) of the worst block for the worst channel
    t = y.double()
    blk = t[:, :720].reshape(69, 45, 16, 90, 16).mean(dim=(2, 4)).cpu().numpy()
    d = np.abs(blk - fx["y_block"]) / fx["y_std"][:, None, None]
    ch = int(d.max(axis=(1, 2)).argmax())
    print("  worst block channel", PANGU_CHANNELS[ch], "lat-block profile (max over lon):", np.round(d[ch].max(axis=1), 4).tolist())
    print("  lon-block profile (max over lat):", np.round(d[ch].max(axis=0), 4).tolist()[:30])
    eng.close(); del y, t
if "sfno" in what:
    from oracle.sfno_ref import SFNORef
    for (nlat, nlon, E, L) in [(97, 192, 128, 2), (145, 288, 64, 1), (145, 288, 128, 2), (193, 384, 64, 1), (241, 480, 64, 1), (241, 480, 128, 2), (241, 96, 64, 1), (49, 480, 64, 1)]:
        try:
            cfg = sfno_small(nlat, nlon, embed=E, layers=L)
            w = make_sfno_weights(cfg, 3)
            x0 = synthetic_state(FCNV2_CHANNELS, nlat, nlon, 1)
            allw = dict(w); allw.update(sfno_tables(cfg))
            eng = StepEngine(cfg, 0); eng.load_weights(allw)
            y = eng.step(torch.from_numpy(x0)[None].cuda())[0].cpu().numpy()
            eng.close()
            e = rel_err_per_channel(y, SFNORef(cfg, w).step(x0).numpy())
            print(f"sfno {nlat}x{nlon} E{E} L{L}: h={cfg.h} w={cfg.w} lmax={cfg.lmax} mmax={cfg.mmax}  max rel err {e.max():.3e}", flush=True)
        except Exception as ex:
            print(f"sfno {nlat}x{nlon} E{E} L{L}: EXC {ex}", flush=True)
