#!/bin/bash
# Round-2 ncu evidence (run under gpurun on ONE GPU; outputs under gpurun_out/, summaries are copied to profiles/ by hand).
#   1. launch list of one prefill step (device time per launch, cold-cache / serialised: compare SHARES)
#   2. DRAM bytes per launch of the Llama-prefill GEMMs of the same step (roofline.traffic)
#   3. ncu --set full capture of the dominant kernel (gate|up weight-streaming GEMM, production configuration)
set -x
OUT=gpurun_out
K='regex:gemm_tc|attn|layernorm|rmsnorm|swiglu|rope|logmel|lm_head|argmax|splice|splitk|mel_to|kv_|add_i32|tile_weight'
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 2400 --csv --log-file $OUT/launches_r2.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-library-baseline --no-train-record --ttft-iters 1 > $OUT/ncu_bench_r2.log 2>&1
timeout -s KILL 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tc -c 1200 --csv \
  --log-file $OUT/dram_gemm_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-library-baseline --no-train-record --ttft-iters 1 > $OUT/ncu_dram_r2.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 6 -c 2 -o $OUT/prof_r2_gate_up \
  python scripts/gemm_one_r2.py gate_up > $OUT/ncu_full_r2.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:attn_llm_tc -s 2 -c 1 -o $OUT/prof_r2_attn_llm \
  python scripts/gemm_one_r2.py attn > $OUT/ncu_full_attn_r2.log 2>&1
ls -la $OUT | tail -8
