"""Real-weight import: the protobuf reader, the bias un-gathering and the two name mappings, on
files this test writes itself with the same wire formats (no onnx / onnxruntime / earth2mip here)."""
import numpy as np
import pytest
import torch

from skyrim_b200 import importers as I
from skyrim_b200.config import pangu_small, sfno_small
from skyrim_b200.weights import make_pangu_weights, make_sfno_weights


def test_protobuf_reader_roundtrip(tmp_path):
    t = {"a": np.arange(24, dtype=np.float32).reshape(2, 3, 4), "b": np.array([3, -1], np.int64),
         "h": np.arange(5, dtype=np.float16), "s": np.float32(2.5).reshape(())}
    p = str(tmp_path / "m.onnx")
    I.write_onnx_initializers(p, t, ops=[("MatMul", ["x", "a"], ["y"]), ("Add", ["y", "b"], ["z"])])
    got, ops = I.read_onnx_initializers(p)
    assert list(got) == list(t) and [o[0] for o in ops] == ["MatMul", "Add"] and ops[0][1] == ["x", "a"]
    for k in t:
        assert got[k].dtype == t[k].dtype and got[k].shape == t[k].shape and np.array_equal(got[k], t[k])


def test_position_index_matches_the_oracle_and_ungather_inverts_the_gather():
    from oracle.pangu_ref import position_index
    idx = I.position_index((2, 6, 12))
    assert np.array_equal(idx, position_index((2, 6, 12)).numpy())
    rng = np.random.default_rng(0)
    table = rng.standard_normal((3312, 5, 6)).astype(np.float32)
    expanded = np.moveaxis(table[idx.reshape(-1)].reshape(144, 144, 5, 6), (2, 3), (0, 1))   # (type, head, i, j)
    assert np.array_equal(I.ungather_bias(expanded), table)
    expanded[2, 3, 7, 9] += 1.0
    with pytest.raises(ValueError):
        I.ungather_bias(expanded)


def test_pangu_from_onnx_on_a_synthetic_export(tmp_path):
    """Weights in the form an exporter leaves them: anonymous names, MatMul weights [in, out], expanded biases."""
    cfg = pangu_small(41, 96)
    w = make_pangu_weights(cfg, 3)
    idx = I.position_index(cfg.window).reshape(-1)
    exported, k = {}, 0
    for name, a in w.items():
        if name in ("norm.mean", "norm.std", "const.masks"):
            continue
        if name.endswith("bias_table"):
            a = np.moveaxis(a[idx].reshape(144, 144, a.shape[1], a.shape[2]), (2, 3), (0, 1))
        elif a.ndim == 2:
            a = a.T
        exported[f"onnx::T_{k}"] = np.ascontiguousarray(a); k += 1
    exported["mean"] = w["norm.mean"]; exported["std"] = w["norm.std"]
    exported["masks"] = w["const.masks"]
    p = str(tmp_path / "pangu_weather_6.onnx")
    I.write_onnx_initializers(p, exported)
    got = I.pangu_from_onnx(p, cfg)
    assert list(got) == list(w)
    for name in w:
        assert np.array_equal(got[name], w[name]), name
    assert I.check_fp16_range(got)[1] < 10.0


def test_sfno_from_checkpoint_on_a_synthetic_checkpoint(tmp_path):
    cfg = sfno_small(49, 96, embed=64, layers=2)
    w = make_sfno_weights(cfg, 1)
    sd = {}
    conv = lambda a: torch.from_numpy(a)[:, :, None, None]
    sd["module.encoder.fwd.0.weight"], sd["module.encoder.fwd.0.bias"] = conv(w["enc.fc1.w"]), torch.from_numpy(w["enc.fc1.b"])
    sd["module.encoder.fwd.2.weight"], sd["module.encoder.fwd.2.bias"] = conv(w["enc.fc2.w"]), torch.from_numpy(w["enc.fc2.b"])
    sd["module.pos_embed"] = torch.from_numpy(w["pos_embed"])[None]
    sd["module.decoder.fwd.0.weight"], sd["module.decoder.fwd.0.bias"] = conv(w["dec.fc1.w"]), torch.from_numpy(w["dec.fc1.b"])
    sd["module.decoder.fwd.2.weight"], sd["module.decoder.fwd.2.bias"] = conv(w["dec.fc2.w"]), torch.from_numpy(w["dec.fc2.b"])
    for i in range(cfg.layers):
        b, p = f"module.blocks.{i}.", f"blk{i}."
        sd[b + "norm0.weight"], sd[b + "norm0.bias"] = torch.from_numpy(w[p + "norm0.g"]), torch.from_numpy(w[p + "norm0.b"])
        sd[b + "norm1.weight"], sd[b + "norm1.bias"] = torch.from_numpy(w[p + "norm1.g"]), torch.from_numpy(w[p + "norm1.b"])
        sw = torch.from_numpy(w[p + "spec.w"])                       # [l, out, in, 2]
        sd[b + "filter.filter.weight"] = torch.view_as_complex(sw.permute(2, 1, 0, 3).contiguous())   # (in, out, l) complex
        sd[b + "inner_skip.weight"], sd[b + "inner_skip.bias"] = conv(w[p + "inner.w"]), torch.from_numpy(w[p + "inner.b"])
        sd[b + "mlp.fwd.0.weight"], sd[b + "mlp.fwd.0.bias"] = conv(w[p + "fc1.w"]), torch.from_numpy(w[p + "fc1.b"])
        sd[b + "mlp.fwd.2.weight"], sd[b + "mlp.fwd.2.bias"] = conv(w[p + "fc2.w"]), torch.from_numpy(w[p + "fc2.b"])
    torch.save({"model_state": sd}, str(tmp_path / "weights.tar"))
    np.save(str(tmp_path / "global_means.npy"), w["norm.mean"].reshape(1, -1, 1, 1))
    np.save(str(tmp_path / "global_stds.npy"), w["norm.std"].reshape(1, -1, 1, 1))
    got_cfg, got = I.load_real_weights("sfno", str(tmp_path))
    assert (got_cfg.nlat, got_cfg.nlon, got_cfg.embed, got_cfg.layers, got_cfg.scale_factor, got_cfg.n_channels) == (49, 96, 64, 2, 3, 73)
    assert list(got) == list(w)
    for name in w:
        assert np.array_equal(got[name], w[name]), name


def test_fp16_range_guard():
    with pytest.raises(ValueError):
        I.check_fp16_range({"layer0.block0.fc1.w": np.array([1.0, 7.0e4], np.float32)})
    assert I.check_fp16_range({"norm.mean": np.array([1e5], np.float32), "x.w": np.array([0.5], np.float32)}) == ("x.w", 0.5)
