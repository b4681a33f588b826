"""N full-size steps of one model through the product library (for ncu captures).  python tools/gpu_steps.py pangu|sfno [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from skyrim_b200.config import *
from skyrim_b200.engine import StepEngine
from skyrim_b200.weights import *
model = sys.argv[1] if len(sys.argv) > 1 else "pangu"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if model == "pangu":
    cfg, ch = pangu_full(), PANGU_CHANNELS
    w = make_pangu_weights(cfg, 0)
else:
    cfg, ch = sfno_full(), FCNV2_CHANNELS
    w = dict(make_sfno_weights(cfg, 0)); w.update(sfno_tables(cfg))
eng = StepEngine(cfg, 0); eng.load_weights(w); del w
x = torch.from_numpy(synthetic_state(ch, cfg.nlat, cfg.nlon, 0))[None].cuda()
y = torch.empty_like(x)
torch.cuda.synchronize()
print("STEPS_BEGIN", flush=True)
for _ in range(n):
    eng.step(x, y); x, y = y, x
torch.cuda.synchronize()
print("done", bool(torch.isfinite(x).all()))
