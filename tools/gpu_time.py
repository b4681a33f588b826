"""Time the full-size Pangu step kernel by kernel-sequence (CUDA events)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from skyrim_b200.config import pangu_full, PANGU_CHANNELS
from skyrim_b200.weights import make_pangu_weights, synthetic_state
from skyrim_b200.engine import StepEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = pangu_full()
t = time.time(); w = make_pangu_weights(cfg, 0); print("weights %.1fs" % (time.time() - t), flush=True)
x0 = synthetic_state(PANGU_CHANNELS, cfg.nlat, cfg.nlon, 0)
eng = StepEngine(cfg, 0); eng.load_weights(w)
x = torch.from_numpy(x0)[None].repeat(B, 1, 1, 1).cuda().contiguous()
y = torch.empty_like(x)
for _ in range(3):
    eng.step(x, y); x, y = y, x
torch.cuda.synchronize()
print("finite:", bool(torch.isfinite(x).all()), "normalised std:", float(((x[0, 39:65] ) / 10).std()), flush=True)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
ev[0].record()
for i in range(10):
    eng.step(x, y); x, y = y, x
    ev[i + 1].record()
torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(10)]
print("B=%d ms/step: %s  -> %.2f member-steps/s" % (B, ["%.2f" % t for t in ts], B * 1000 / np.median(ts)), flush=True)
# per-family breakdown (events around every launch)
eng.profile_begin()
for i in range(3):
    eng.step(x, y); x, y = y, x
prof = eng.profile_end()
tot = sum(v[0] for v in prof.values())
for k, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print("  %-8s %8.3f ms/step  %3d launches/step  %5.1f%%" % (k, ms / 3, n // 3, 100 * ms / tot), flush=True)
print("  sum %.2f ms/step" % (tot / 3))
