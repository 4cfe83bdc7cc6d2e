timeout 600 python tools/dag_sweep.py 512 1024 2048 4096 8192 12288 16384 24576 32768 > gpurun_out/r02_size_sweep.jsonl 2>&1
for n in 200 330 500 1000; do timeout 120 python tools/mid_batch_bench.py $n 256; GPEDM_NO_MID_BATCH=1 timeout 120 python tools/mid_batch_bench.py $n 256; done > gpurun_out/r02_mid_batch.jsonl 2>&1
timeout 120 python tools/small_batch_bench.py 592 > gpurun_out/r02_small_batch.jsonl 2>&1
timeout 900 python tools/run_configs.py C1 C2 C3 C5 > gpurun_out/r02_configs_gpu.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_bench_v3.csv python bench.py --steps 2 --warmup 3 --no-grid --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dag_factor -s 2 -c 1 -o gpurun_out/r02_dag_v3 -f python tools/one_eval.py 8192 3 > gpurun_out/r02_ncu_dag.log 2>&1
ncu --set full --clock-control none -k regex:"assemble_sigma|grad_trace" -s 4 -c 2 -o gpurun_out/r02_hbm_v3 -f python tools/one_eval.py 8192 3 > gpurun_out/r02_ncu_hbm.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_dag_v3.ncu-rep > gpurun_out/r02_dag_v3_summary.txt
python tools/ncu_summary.py gpurun_out/r02_hbm_v3.ncu-rep > gpurun_out/r02_hbm_v3_summary.txt
timeout 200 python tools/dag_timeline.py 8192 500 > gpurun_out/r02_dag_tl_8192_v3.txt 2>&1
timeout 200 python tools/dag_timeline.py 4096 100 > gpurun_out/r02_dag_tl_4096_v3.txt 2>&1
timeout 200 python tools/dag_timeline.py 1024 50 > gpurun_out/r02_dag_tl_1024_v3.txt 2>&1
tail -3 gpurun_out/r02_size_sweep.jsonl; cat gpurun_out/r02_mid_batch.jsonl
