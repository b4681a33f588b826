"""Perturbed-IC ensemble on one GPU (the multi-GPU partition is covered on CPU by test_dist_cpu)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_ensemble_runner_members_and_spread():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from skyrim_b200.config import PANGU_CHANNELS, pangu_small
    from skyrim_b200.engine import StepEngine
    from skyrim_b200.ensemble import EnsembleRunner
    from skyrim_b200.weights import channel_stats, make_pangu_weights, synthetic_state
    cfg = pangu_small(41, 96)
    eng = StepEngine(cfg, 0)
    eng.load_weights(make_pangu_weights(cfg, 0))
    base = synthetic_state(PANGU_CHANNELS, cfg.nlat, cfg.nlon, 0)
    sig = channel_stats(PANGU_CHANNELS)[1]
    run = EnsembleRunner(eng, base, sig, members_per_gpu=4, rank=0, amp=0.05, seed=3)
    x0 = run.x.clone()
    # member m of a 4-member rank equals the same global member computed by a 1-member "rank"
    solo = EnsembleRunner(eng, base, sig, members_per_gpu=1, rank=2, amp=0.05, seed=3)
    assert torch.equal(solo.x[0], x0[2])
    run.step(2); solo.step(2)
    assert torch.equal(solo.x[0], run.x[2])
    mean, spread = run.mean_and_spread(world=1)
    assert mean.shape == (69, 41, 96) and bool(torch.isfinite(spread).all()) and float(spread.mean()) > 0
    eng.close()
