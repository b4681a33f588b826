"""GPU parity of the SFNO step (CUDA engine through the C-ABI) against the CPU oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.mark.parametrize("nlat,nlon,embed,layers", [(49, 96, 64, 3), (97, 192, 128, 2)])
def test_sfno_step_parity(nlat, nlon, embed, layers):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle.pangu_ref import rel_err_per_channel
    from oracle.sfno_ref import SFNORef
    from skyrim_b200.config import FCNV2_CHANNELS, sfno_small
    from skyrim_b200.engine import StepEngine
    from skyrim_b200.weights import make_sfno_weights, sfno_tables, synthetic_state
    cfg = sfno_small(nlat, nlon, embed=embed, layers=layers)
    w = make_sfno_weights(cfg, 0)
    x0 = synthetic_state(FCNV2_CHANNELS, cfg.nlat, cfg.nlon, 0)
    eng = StepEngine(cfg, 0)
    allw = dict(w); allw.update(sfno_tables(cfg))
    eng.load_weights(allw)
    y = eng.step(torch.from_numpy(x0)[None].cuda())[0].cpu().numpy()
    ref = SFNORef(cfg, w).step(x0).numpy()
    e = rel_err_per_channel(y, ref)
    assert np.isfinite(y).all() and e.max() < TOL, (e.max(), FCNV2_CHANNELS[int(e.argmax())])
    # two chained steps through the TimeLoop protocol
    from skyrim_b200.timeloop import SFNOTimeLoop
    import datetime
    loop = SFNOTimeLoop(eng)
    it = loop(datetime.datetime(2024, 5, 7), torch.from_numpy(x0)[None, None])
    next(it); next(it)
    _, y2, _ = next(it)
    ref2 = SFNORef(cfg, w).step(ref).numpy()
    assert rel_err_per_channel(y2[0].cpu().numpy(), ref2).max() < 2 * TOL
    eng.close()


def test_sfno_step_against_golden_fixture():
    """CUDA engine against the committed fp64 fixture (tests/golden/sfno_49x96_seed0.npz, tools/make_golden.py):
    sampled outputs and per-channel norms, no oracle evaluation in the loop."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import os
    from skyrim_b200.config import FCNV2_CHANNELS, sfno_small
    from skyrim_b200.engine import StepEngine
    from skyrim_b200.weights import make_sfno_weights, sfno_tables, synthetic_state
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sfno_49x96_seed0.npz"))
    cfg = sfno_small(49, 96, embed=64, layers=3)
    w = make_sfno_weights(cfg, 0)
    x0 = synthetic_state(FCNV2_CHANNELS, cfg.nlat, cfg.nlon, 0)
    eng = StepEngine(cfg, 0)
    allw = dict(w); allw.update(sfno_tables(cfg))
    eng.load_weights(allw)
    y = eng.step(torch.from_numpy(x0)[None].cuda())[0].cpu().numpy()
    eng.close()
    scale = np.abs(g["y_sample"]).max(axis=(1, 2), keepdims=True)
    assert np.max(np.abs(y[:, ::6, ::12] - g["y_sample"]) / scale) < TOL
    np.testing.assert_allclose(np.sqrt((y.astype(np.float64) ** 2).sum(axis=(1, 2))), g["y_norm"], rtol=TOL)
