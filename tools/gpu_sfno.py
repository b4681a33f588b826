"""SFNO step on the GPU vs the oracle (small) and timing at full size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from skyrim_b200.config import sfno_small, sfno_full, FCNV2_CHANNELS
from skyrim_b200.weights import make_sfno_weights, sfno_tables, synthetic_state
from skyrim_b200.engine import StepEngine
from oracle.sfno_ref import SFNORef
from oracle.pangu_ref import rel_err_per_channel
mode = sys.argv[1] if len(sys.argv) > 1 else "small"
if mode == "small":
    for (a, b, e, l) in [(49, 96, 64, 3), (97, 192, 128, 2)]:
        cfg = sfno_small(a, b, embed=e, layers=l)
        w = make_sfno_weights(cfg, 0); x0 = synthetic_state(FCNV2_CHANNELS, cfg.nlat, cfg.nlon, 0)
        eng = StepEngine(cfg, 0); allw = dict(w); allw.update(sfno_tables(cfg)); eng.load_weights(allw)
        y = eng.step(torch.from_numpy(x0)[None].cuda())[0].cpu().numpy()
        ref = SFNORef(cfg, w).step(x0).numpy()
        err = rel_err_per_channel(y, ref)
        print(f"SFNO {a}x{b} E={e} L={l}: finite={np.isfinite(y).all()} rel err max {err.max():.3e} median {np.median(err):.3e}", flush=True)
        eng.close()
else:
    cfg = sfno_full()
    t = time.time(); w = make_sfno_weights(cfg, 0); w.update(sfno_tables(cfg)); print("weights+tables %.1fs" % (time.time() - t), flush=True)
    x0 = synthetic_state(FCNV2_CHANNELS, cfg.nlat, cfg.nlon, 0)
    eng = StepEngine(cfg, 0); t = time.time(); eng.load_weights(w); torch.cuda.synchronize(); print("load %.1fs" % (time.time() - t), flush=True)
    del w
    x = torch.from_numpy(x0)[None].cuda(); y = torch.empty_like(x)
    for _ in range(2):
        eng.step(x, y); x, y = y, x
    torch.cuda.synchronize()
    print("finite", bool(torch.isfinite(x).all()), flush=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    for i in range(5):
        eng.step(x, y); x, y = y, x; ev[i + 1].record()
    torch.cuda.synchronize()
    print("ms/step", ["%.1f" % ev[i].elapsed_time(ev[i + 1]) for i in range(5)], flush=True)
    eng.profile_begin(); eng.step(x, y); prof = eng.profile_end()
    for k, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        print("  %-10s %8.2f ms  %3d launches" % (k, ms, n), flush=True)
