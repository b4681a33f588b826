"""Two small-grid Pangu steps and one SFNO step (target for compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from skyrim_b200.config import pangu_small, sfno_small, PANGU_CHANNELS, FCNV2_CHANNELS
from skyrim_b200.weights import make_pangu_weights, make_sfno_weights, sfno_tables, synthetic_state
from skyrim_b200.engine import StepEngine

cfg = pangu_small(41, 96)
eng = StepEngine(cfg, 0); eng.load_weights(make_pangu_weights(cfg, 0))
x = torch.from_numpy(synthetic_state(PANGU_CHANNELS, cfg.nlat, cfg.nlon, 0))[None].repeat(2, 1, 1, 1).cuda()
y = torch.empty_like(x)
for _ in range(2):
    eng.step(x, y); x, y = y, x
torch.cuda.synchronize()
print("pangu done", float(x.mean()), bool(torch.isfinite(x).all()))
eng.close()
if "--sfno" in sys.argv:
    c2 = sfno_small()
    w2 = dict(make_sfno_weights(c2, 0)); w2.update(sfno_tables(c2))
    e2 = StepEngine(c2, 0); e2.load_weights(w2)
    xs = torch.from_numpy(synthetic_state(FCNV2_CHANNELS, c2.nlat, c2.nlon, 0))[None].cuda()
    ys = e2.step(xs); torch.cuda.synchronize()
    print("sfno done", float(ys.mean()), bool(torch.isfinite(ys).all()))
