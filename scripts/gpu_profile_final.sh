#!/bin/bash
# Round-2 evidence of the final build: ncu launch list of one step, ncu --set full of the dominant kernels, bench line
# with per-shape tables, reference arm.
mkdir -p gpurun_out
SPGAN_BENCH_MINIMAL=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 200 $NCU -k regex:gemm_ts_kernel -s 2 -c 1 -o gpurun_out/r2f_ts_fc2 python scripts/prof_fused_shape.py 131072 1024 256 pro > gpurun_out/ncu_a.log 2>&1
timeout 200 $NCU -k regex:gemm_ts_kernel -s 2 -c 1 -o gpurun_out/r2f_ts_e128k64 python scripts/prof_fused_shape.py 1310720 128 64 pro > gpurun_out/ncu_b.log 2>&1
timeout 200 $NCU -k regex:gemm_ts_kernel -s 2 -c 1 -o gpurun_out/r2f_ts_n256k128 python scripts/prof_fused_shape.py 131072 256 128 pro > gpurun_out/ncu_b2.log 2>&1
timeout 200 $NCU -k regex:gemm_wg_kernel -s 2 -c 1 -o gpurun_out/r2f_wg_fc2 python scripts/prof_wgrad_shape.py 1024 256 131072 > gpurun_out/ncu_c.log 2>&1
timeout 200 $NCU -k regex:gemm_tsk_kernel -s 2 -c 1 -o gpurun_out/r2f_tsk_k1280 python scripts/prof_gemm_shape.py 131072 128 1280 > gpurun_out/ncu_d.log 2>&1
timeout 200 $NCU -k regex:gemm_tsk_kernel -s 2 -c 1 -o gpurun_out/r2f_tsk_k1024 python scripts/prof_gemm_shape.py 131072 256 1024 > gpurun_out/ncu_d2.log 2>&1
timeout 200 $NCU -k regex:"knn_tc_filter|knn_tc_refine" -s 2 -c 2 -o gpurun_out/r2f_knn_tc python scripts/prof_knn_tc.py 64 > gpurun_out/ncu_e.log 2>&1
timeout 300 $NCU -k regex:"edge_combine_stats|attn_bwd_stats|edge_combine_bwd_bn|bn_softmax_mul_k_kernel|colmap4_kernel|colreduce4_kernel|bn_pool_partial" -s 12 -c 14 -o gpurun_out/r2f_edge python scripts/prof_edgeblock.py > gpurun_out/ncu_f.log 2>&1
SPGAN_BENCH_BW_TABLE=1 SPGAN_BENCH_GEMM_TABLE=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err
ls -la gpurun_out | grep r2f; cat gpurun_out/ncu_f.log | tail -3; tail -c 400 gpurun_out/bench_r2_reference.json; tail -c 300 gpurun_out/bench_r2_final.json
