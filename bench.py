#!/usr/bin/env python
"""Benchmark of the rollout hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 3            # our arm, 1 GPU
    torchrun --nproc-per-node N ... bench.py --gpus N ...     # one rank per GPU, weak scaling
    python bench.py --impl reference --steps 5 --warmup 1     # CPU reference arm (the oracle)

A "step" is one 6-h Pangu step of every member resident on the GPU (``--members-per-gpu``,
default 1 = config[1] "Pangu 7-day rollout, synthetic IC, 1xB200": the chained device-resident
rollout).  ``value`` = member-steps per second over all ranks with the state resident in HBM;
``e2e`` = the same through the reference-facing TimeLoop call with HOST (pinned) input and output
every step, copies inside the timed region.  One JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "6h rollout steps/sec on (69,721,1440); ensemble member-steps/sec @1/2/4/8 GPU"
UNIT = "member-steps/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"],
                    tensor_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle (restatement of the reference's forward) on the
# host cores, on a bounded latitude band of the same workload, scaled by the FLOP ratio.
# --------------------------------------------------------------------------------------------
def cpu_reference(steps: int, warmup: int, band_nlat: int):
    import numpy as np
    import torch
    from oracle.pangu_ref import PanguRef
    from skyrim_b200.config import PANGU_CHANNELS, pangu_full, pangu_small
    from skyrim_b200.roofline import pangu_flops
    from skyrim_b200.weights import make_pangu_weights, synthetic_state
    # torch's CPU kernels stop scaling (and oversubscribe) beyond a few dozen threads on the
    # big bench hosts; `cores` reports the threads actually used
    cores = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(cores)
    full, band = pangu_full(), pangu_small(band_nlat, 1440)
    frac = pangu_flops(band)["total"] / pangu_flops(full)["total"]
    w = make_pangu_weights(band, 0)
    x = torch.from_numpy(synthetic_state(PANGU_CHANNELS, band.nlat, band.nlon, 0))
    ref = PanguRef(band, w)
    for _ in range(warmup):
        ref.step(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        x = ref.step(x)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    assert bool(torch.isfinite(x).all())
    return dict(value=frac / dt, unit=UNIT, cores=cores, kind="port", sec_per_sample=dt,
                sample=(f"oracle (torch fp32, {cores} threads) on a {band_nlat}x1440 latitude band = "
                        f"{100 * frac:.1f}% of the full step's FLOPs; value = band fraction / seconds"))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    band = 49 if total <= 30 else 25
    cb = cpu_reference(args.steps, args.warmup, band)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / cb["value"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Pangu 6-h step, synthetic (69,721,1440) IC, CPU restatement of the reference "
                                   "forward (onnxruntime / earth2mip are not installable offline)"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from skyrim_b200.config import PANGU_CHANNELS, pangu_full
    from skyrim_b200.engine import StepEngine, launch_count, pack_arena, perturb_ic
    from skyrim_b200.roofline import pangu_flops, pangu_state_bytes
    from skyrim_b200.timeloop import PanguTimeLoop
    from skyrim_b200.weights import channel_stats, make_pangu_weights, synthetic_state

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sfno = args.model == "sfno"
    if sfno:
        from skyrim_b200.config import FCNV2_CHANNELS as CHANNELS, sfno_full
        from skyrim_b200.roofline import sfno_flops
        from skyrim_b200.timeloop import SFNOTimeLoop as Loop
        from skyrim_b200.weights import make_sfno_weights, sfno_param_shapes, sfno_tables
        cfg = sfno_full()
    else:
        CHANNELS, Loop = PANGU_CHANNELS, PanguTimeLoop
        cfg = pangu_full()
    M = args.members_per_gpu
    # ---- weights: built on rank 0, ONE NCCL broadcast of the fp32 arena, repacked on each device ----
    if sfno:
        w = None
        if rank == 0:
            w = make_sfno_weights(cfg, 0); w.update(sfno_tables(cfg))
    else:
        w = make_pangu_weights(cfg, 0) if rank == 0 else None
    if world > 1:
        from skyrim_b200.weights import pangu_param_shapes
        if sfno:
            shapes = dict(sfno_param_shapes(cfg))
            shapes.update({k: v.shape for k, v in sfno_tables(cfg).items()})
        else:
            shapes = pangu_param_shapes(cfg)
        if rank == 0:
            arena_h, manifest = pack_arena(w)
            arena = torch.from_numpy(arena_h).to(dev)
        else:
            arena_h, manifest = pack_arena({k: np.zeros(s, np.float32) for k, s in shapes.items()})
            arena = torch.empty(arena_h.size, dtype=torch.float32, device=dev)
        dist.broadcast(arena, 0)
        eng = StepEngine(cfg, local)
        eng.load_arena(arena, manifest)
        del arena
    else:
        eng = StepEngine(cfg, local)
        eng.load_weights(w)
    loop = Loop(eng)
    del w

    # ---- synthetic initial conditions: base state + per-member Philox perturbation (K11) ----
    base = torch.from_numpy(synthetic_state(CHANNELS, cfg.nlat, cfg.nlon, 0))
    x = base[None].repeat(M, 1, 1, 1).to(dev).contiguous()
    sigma = torch.from_numpy(channel_stats(CHANNELS)[1]).to(dev)
    perturb_ic(x, sigma, 0.05, seed=0, member0=rank * M)
    y = torch.empty_like(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up + family breakdown (all families timed, outside the timed region) ----
    for _ in range(max(args.warmup - 1, 0)):
        eng.step(x, y); x, y = y, x
    eng.profile_begin()
    eng.step(x, y); x, y = y, x
    fam = eng.profile_end()
    dominant = max(fam, key=lambda k: fam[k][0])

    # ---- timed region: K device-resident chained steps ----
    sampler = ClockSampler(local)
    launches0 = launch_count()
    barrier()
    sampler.start()
    eng.profile_begin([dominant])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.step(x, y); x, y = y, x
    e1.record()
    barrier()
    clocks = sampler.stop()
    dom = eng.profile_end()[dominant]
    ms_total = e0.elapsed_time(e1)
    launches = launch_count() - launches0
    finite = bool(torch.isfinite(x).all())

    # ---- end to end through the TimeLoop call: host (pinned) in, host (pinned) out, every step ----
    xh = torch.empty((M,) + tuple(base.shape), dtype=torch.float32).pin_memory()
    xh.copy_(x.cpu())
    e2e_steps = max(3, min(args.steps, 10))
    out_h = loop.step_host(xh)  # warm-up (allocates the pinned result buffer)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out_h = loop.step_host(out_h)  # H2D + step + D2H, synchronous result on the host
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1000.0 / e2e_steps

    t = torch.tensor([ms_total, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    ms_step = ms_total / args.steps
    value = world * M * 1000.0 / ms_step

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not sfno:
        cb = cpu_reference(1, 0, 91)

    if rank == 0:
        peaks = _peaks()
        fl = sfno_flops(cfg) if sfno else pangu_flops(cfg)
        fam_flops = M * fl.get(dominant, 0.0)
        n_l = dom[1] / args.steps
        avg_ms = dom[0] / max(dom[1], 1)
        achieved = fam_flops / max(n_l, 1) / (avg_ms * 1e-3) / 1e12 if fam_flops else None
        sbytes = cfg.n_channels * cfg.nlat * cfg.nlon * 4 * M
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": ("f16 hi+lo split operands (3-term) / f32 accumulate (tcgen05 kind::f16), f32 state" if sfno else
                      "f16 operands / f32 accumulate (tcgen05 kind::f16), f32 state, LN, softmax"),
            "data": "synthetic",
            "config": {"workload": ("FourCastNet-v2 SFNO rollout (chained 6-h steps), 73-channel synthetic (73,721,1440) IC, "
                                    "state resident in HBM" if sfno else
                                    "Pangu 7-day rollout (chained 6-h steps), synthetic (69,721,1440) IC, state "
                                    "resident in HBM"), "members_per_gpu": M, "members_total": world * M,
                       "l2": "inputs larger than L2 (state 0.3 GB + >2 GB activations streamed per step)",
                       "weights": "synthetic seed 0", "finite": finite},
            "e2e": {"value": world * M * 1000.0 / e2e_ms, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": sbytes, "d2h_bytes_per_step": sbytes,
                    "path": "TimeLoop.step_host: pinned host -> HBM, sky_model_step, HBM -> pinned host"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": dominant, "achieved": achieved,
                         "peak": peaks["tensor_sustained"], "unit": "TFLOP/s",
                         "frac": (achieved / peaks["tensor_sustained"]) if achieved else None, "traffic": _traffic(dominant),
                         "peak_source": peaks["source"] + " bf16 cuBLAS, sustained (kernel timed inside a long step)",
                         "launches_per_step": n_l, "avg_launch_ms": avg_ms,
                         "step_tflops": M * fl["total"] / (ms_step * 1e-3) / 1e12},
            "families_ms_per_step": {k: round(v[0], 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])},
            "cpu_baseline": ({k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")} if cb else None),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _traffic(family):
    """DRAM bytes per launch of the dominant kernel family from the committed `ncu --set full` capture
    (profiles/r1_traffic.json, provenance inside); None when no capture covers the family."""
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get(family, {}).get("bytes_per_launch")
    except (OSError, ValueError):
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=28)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--members-per-gpu", type=int, default=1)
    ap.add_argument("--model", default="pangu", choices=["pangu", "sfno"],
                    help="pangu (default, the headline workload) or sfno (FourCastNet-v2 73-channel rollout)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
