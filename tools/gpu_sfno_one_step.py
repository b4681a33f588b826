"""Run a couple of full-size SFNO steps (target for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from skyrim_b200.config import sfno_full, FCNV2_CHANNELS
from skyrim_b200.weights import make_sfno_weights, sfno_tables, synthetic_state
from skyrim_b200.engine import StepEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = sfno_full()
w = make_sfno_weights(cfg, 0); w.update(sfno_tables(cfg))
eng = StepEngine(cfg, 0); eng.load_weights(w); del w
x = torch.from_numpy(synthetic_state(FCNV2_CHANNELS, cfg.nlat, cfg.nlon, 0))[None].cuda()
y = torch.empty_like(x)
for _ in range(n):
    eng.step(x, y); x, y = y, x
torch.cuda.synchronize()
print("done", float(x.mean()))
