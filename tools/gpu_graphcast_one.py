"""one full-size GraphCast step after load (for ncu launch lists / captures): python tools/gpu_graphcast_one.py [n_steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from skyrim_b200 import icomesh
from skyrim_b200.config import graphcast_full
from skyrim_b200.engine import StepEngine
from skyrim_b200.timeloop import GraphcastTimeLoop
from skyrim_b200.weights import make_graphcast_weights, synthetic_graphcast_state
cfg = graphcast_full()
eng = StepEngine(cfg, 0)
eng.load_weights(make_graphcast_weights(cfg, 0))
eng.debug_set("use_graphs", 0)
loop = GraphcastTimeLoop(eng)
x = torch.from_numpy(synthetic_graphcast_state(cfg, 0)).reshape(1, 2, cfg.n_state, cfg.nlat, cfg.nlon).cuda()
loop.fill_forcing(x, 1714521600.0)
x = x.reshape(1, -1, cfg.nlat, cfg.nlon).contiguous()
y = torch.empty_like(x)
eng.set_clock(1714521600.0)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    eng.step(x, y); x, y = y, x
torch.cuda.synchronize()
print("done")
