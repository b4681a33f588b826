"""Skyrim(name).predict(...) end to end on the GPU engine (small grid)."""
import datetime

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_pangu_predict_api(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle.pangu_ref import PanguRef, rel_err_per_channel
    from skyrim_b200 import Skyrim
    from skyrim_b200.config import PANGU_CHANNELS, pangu_small
    from skyrim_b200.weights import make_pangu_weights
    cfg = pangu_small(41, 96)
    w = make_pangu_weights(cfg, 0)
    sk = Skyrim("pangu", ic_source="synthetic", cfg=cfg, weights=w)
    pred, paths = sk.predict(date="20240507", time="0000", lead_time=13, save=True,
                             save_config=dict(output_dir=str(tmp_path), file_type="netcdf"))
    # reference semantics: lead time floored to 12 h -> 2 steps, one file per step, dims time/channel/lat/lon
    assert len(paths) == 2 and set(pred.prediction.dims) == {"time", "channel", "lat", "lon"}
    assert list(pred.channels.values) == PANGU_CHANNELS
    assert pred.prediction.shape == (2, 69, 41, 96)
    ref = PanguRef(cfg, w)
    x0 = sk.model.data_source[datetime.datetime(2024, 5, 7)].values
    y2 = ref.step(ref.step(x0)).numpy()
    e = rel_err_per_channel(pred.prediction.values[1], y2)
    assert e.max() < 2e-3, e.max()
    assert np.isfinite(pred.wind_speed(lat=10.0, lon=20.0, pressure_level=850))
    # forecast(): all steps in one device-resident generator, first slice is the IC
    da = sk.forecast(datetime.datetime(2024, 5, 7), n_steps=2, channels=["t2m", "u10m"])
    assert da.shape == (3, 2, 41, 96)
    np.testing.assert_allclose(da.values[0, 0], x0[PANGU_CHANNELS.index("t2m")])


def test_iter_host_matches_device_resident_chain():
    """Overlapped host delivery (copy of step n under step n+1, ring of three device states, graph replay from the third
    use of a pair) returns bit-identical states to a plain chain of steps."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from skyrim_b200.config import PANGU_CHANNELS, pangu_small
    from skyrim_b200.engine import StepEngine
    from skyrim_b200.timeloop import PanguTimeLoop
    from skyrim_b200.weights import make_pangu_weights, synthetic_state
    cfg = pangu_small(41, 96)
    eng = StepEngine(cfg, 0)
    eng.load_weights(make_pangu_weights(cfg, 0))
    x0 = torch.from_numpy(synthetic_state(PANGU_CHANNELS, cfg.nlat, cfg.nlon, 0))[None]
    loop = PanguTimeLoop(eng)
    t0 = datetime.datetime(2024, 5, 7)
    want, cur = [x0.clone()], x0.cuda()
    for _ in range(7):
        cur = eng.step(cur)
        want.append(cur.cpu())
    for rep in range(2):   # second pass runs on the replayed graphs
        got = [(t, h.clone()) for t, h in loop.iter_host(t0, x0[:, None], 7)]
        assert len(got) == 8 and got[3][0] == t0 + 3 * loop.time_step
        for (t, h), w in zip(got, want):
            assert torch.equal(h, w)
    eng.close()


@pytest.mark.parametrize("model", ["pangu", "sfno"])
def test_fp16_range_guard(model):
    """The guarded step reports max |value| of every fp16 operand image class (O(1) on synthetic weights) and refuses
    weights that push an operand past the fp16 range."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from skyrim_b200 import _ffi
    from skyrim_b200.config import FCNV2_CHANNELS, PANGU_CHANNELS, pangu_small, sfno_small
    from skyrim_b200.engine import StepEngine
    from skyrim_b200.weights import make_pangu_weights, make_sfno_weights, sfno_tables, synthetic_state
    if model == "pangu":
        cfg, ch = pangu_small(41, 96), PANGU_CHANNELS
        w = make_pangu_weights(cfg, 0)
        big = "layer0.block0.qkv.w"
    else:
        cfg, ch = sfno_small(49, 96, embed=64, layers=3), FCNV2_CHANNELS
        w = make_sfno_weights(cfg, 0)
        big = "blk1.fc1.w"
    assert big in w, sorted(w)[:20]
    x = torch.from_numpy(synthetic_state(ch, cfg.nlat, cfg.nlon, 0))[None].cuda()

    def engine(weights):
        eng = StepEngine(cfg, 0)
        allw = dict(weights)
        if model == "sfno":
            allw.update(sfno_tables(cfg))
        eng.load_weights(allw)
        return eng

    eng = engine(w)
    y_plain = eng.step(x).clone()
    y, ranges = eng.step_guarded(x)
    assert torch.equal(y, y_plain)                       # the guard only observes
    assert ranges and all(0.0 < v < 1.0e3 for v in ranges.values()), ranges
    eng.close()
    w2 = dict(w)
    w2[big] = w[big] * 1.0e6                              # one projection far outside what fp16 operands can hold
    eng = engine(w2)
    with pytest.raises(_ffi.SkyError, match="fp16 operand range"):
        eng.step_guarded(x)
    eng.close()
