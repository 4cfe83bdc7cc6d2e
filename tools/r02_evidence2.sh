#!/bin/bash
# Refresh of the round-2 evidence after the trace-kernel change (one B200): bench both arms, ncu launch list, HBM-class
# kernel captures, size sweep, prediction A/B.
O=gpurun_out
python bench.py --steps 10 --warmup 3 > $O/r02g_bench.json 2> $O/r02g_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02g_bench_reference.json 2> $O/r02g_bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02g_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-grid --no-cpu-baseline > $O/r02g_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:assemble_sigma|grad_trace' -s 4 -c 2 -o $O/r02g_hbm -f \
  python tools/one_eval.py 8192 3 > $O/r02g_ncu_hbm.log 2>&1; echo "ncu hbm rc=$?"
timeout 600 python tools/dag_sweep.py 512 1024 2048 4096 8192 12288 16384 24576 32768 > $O/r02g_size_sweep.jsonl 2>&1; echo "sweep rc=$?"
for M in 1 25 1024 4096; do
  GPEDM_LIB=build/libgpedm_prev.so timeout 200 python tools/predict_bench.py 4096 $M | sed 's/^{/{"build": "before (cudaMalloc + cudaFree per call)", /'
  timeout 200 python tools/predict_bench.py 4096 $M | sed 's/^{/{"build": "final (stream-ordered pool)", /'
done > $O/r02g_predict_bench.jsonl 2>&1
timeout 200 python tools/predict_bench.py 8192 4096 | sed 's/^{/{"build": "final (stream-ordered pool)", /' >> $O/r02g_predict_bench.jsonl
timeout 300 python tools/mid_batch_bench.py 330 256 > $O/r02g_mid.jsonl; timeout 300 python tools/mid_batch_bench.py 1000 256 >> $O/r02g_mid.jsonl; timeout 300 python tools/mid_batch_bench.py 4100 24 >> $O/r02g_mid.jsonl
echo done
