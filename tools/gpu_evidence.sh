#!/bin/bash
# Round-end evidence on one B200 (run through gpurun): the bench line, the ncu launch lists of the bench command, and
# `ncu --set full` captures of the dominant kernels.  Everything lands in gpurun_out/ (gpurun brings back at most 64 MiB:
# keep the number of captured launches small); the summaries are copied to profiles/.
tag=${1:-r2}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench_final.json 2> gpurun_out/${tag}_bench_final.err
echo "bench rc=$?"; tail -c 600 gpurun_out/${tag}_bench_final.json
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 560 --csv --log-file gpurun_out/${tag}_launches_pangu.csv \
    python bench.py --steps 3 --warmup 3 --only-headline --no-cpu-baseline > gpurun_out/ncu_lp.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 520 --csv --log-file gpurun_out/${tag}_launches_sfno.csv \
    python bench.py --model sfno --steps 3 --warmup 3 --only-headline --no-cpu-baseline > gpurun_out/ncu_ls.log 2>&1
$NCU --set full --import-source on -k regex:k_mlp_fused_pair --launch-skip 1 -c 2 -f -o gpurun_out/${tag}_mlp_full \
    python tools/gpu_one_step.py 1 > gpurun_out/ncu_m.log 2>&1
$NCU --set full --import-source on -k regex:k_window_attention_tc --launch-skip 1 -c 2 -f -o gpurun_out/${tag}_attn_full \
    python tools/gpu_one_step.py 1 > gpurun_out/ncu_a.log 2>&1
$NCU --set full -k regex:k_gemm_tb -c 5 -f -o gpurun_out/${tag}_sfno_tb_full \
    python tools/gpu_sfno_one_step.py 1 > gpurun_out/ncu_s1.log 2>&1
$NCU --set full -k regex:k_gemm_batched -c 4 -f -o gpurun_out/${tag}_sfno_gb_full \
    python tools/gpu_sfno_one_step.py 1 > gpurun_out/ncu_s2.log 2>&1
$NCU --set full -k regex:k_gemm_batched --launch-skip 48 -c 3 -f -o gpurun_out/${tag}_sfno_gb2_full \
    python tools/gpu_sfno_one_step.py 1 > gpurun_out/ncu_s3.log 2>&1
# ---- GraphCast (config 4) and the kernels added late in round 2 (what produced profiles/r2e_* and r2f_*) ----
$NCU --metrics gpu__time_duration.sum -k regex:"k_gemm|k_gc_" -c 400 --csv --log-file gpurun_out/${tag}_launches_graphcast.csv \
    python tools/gpu_graphcast_one.py 2 > gpurun_out/ncu_lg.log 2>&1
$NCU --set full --import-source on -k regex:k_gemm_split --launch-skip 8 -c 1 -f -o gpurun_out/${tag}_gc_ln_edge \
    python tools/gpu_graphcast_one.py 1 > gpurun_out/ncu_g1.log 2>&1      # mesh-edge LayerNorm GEMM (column-split pair)
$NCU --set full --import-source on -k regex:k_gemm_pair --launch-skip 2 -c 1 -f -o gpurun_out/${tag}_gc_hidden_edge \
    python tools/gpu_graphcast_one.py 1 > gpurun_out/ncu_g2.log 2>&1      # mesh-edge hidden GEMM (A-stationary pair, gathers)
$NCU --set full --import-source on -k regex:k_gemm_split --launch-skip 2 -c 1 -f -o gpurun_out/${tag}_pangu_proj_split \
    python tools/gpu_one_step.py 1 > gpurun_out/ncu_p.log 2>&1           # Pangu C=384 projection (column-split pair)
python tools/gpu_graphcast.py 6 > gpurun_out/${tag}_graphcast_families.log 2>&1
du -sh gpurun_out; ls -la gpurun_out | tail -20
