#!/bin/bash
# Round-2 evidence on the final tree (one B200): tests, bench (both arms), ncu launch list of the bench command,
# --set full captures of the persistent factorisation kernel and of the assembly / trace kernels, size sweep,
# batch rates, BASELINE configs, task timelines.  Everything lands in gpurun_out/r02f_*.
O=gpurun_out
(time python -m pytest tests -m gpu -x -q) > $O/r02f_tests.log 2>&1; tail -4 $O/r02f_tests.log
python bench.py --steps 10 --warmup 3 > $O/r02f_bench.json 2> $O/r02f_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02f_bench_reference.json 2> $O/r02f_bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02f_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-grid --no-cpu-baseline > $O/r02f_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dag_factor -s 2 -c 1 -o $O/r02f_dag -f \
  python tools/one_eval.py 8192 3 > $O/r02f_ncu_dag.log 2>&1; echo "ncu dag rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:assemble_sigma|grad_trace' -s 4 -c 2 -o $O/r02f_hbm -f \
  python tools/one_eval.py 8192 3 > $O/r02f_ncu_hbm.log 2>&1; echo "ncu hbm rc=$?"
timeout 600 python tools/dag_sweep.py 512 1024 2048 4096 8192 12288 16384 24576 32768 > $O/r02f_size_sweep.jsonl 2>&1; echo "sweep rc=$?"
for n in 200 330 500 1000 2050 4100; do
  nf=$((n<1500?256:(n<3000?48:24)))
  timeout 300 python tools/mid_batch_bench.py $n $nf
  GPEDM_NO_MID_BATCH=1 timeout 300 python tools/mid_batch_bench.py $n $nf | sed 's/"workers": "default"/"workers": "default, GPEDM_NO_MID_BATCH=1"/'
done > $O/r02f_mid_batch.jsonl 2>&1; echo "mid rc=$?"
timeout 200 python tools/small_batch_bench.py 592 > $O/r02f_small_batch.jsonl 2>&1
for n in 1024 2048 4096; do timeout 200 python tools/grid_inproc.py 1 $n | tail -1; GPEDM_NO_MID_BATCH=1 timeout 200 python tools/grid_inproc.py 1 $n | tail -1 | sed 's/"workers": "default"/"workers": "default, GPEDM_NO_MID_BATCH=1"/'; done > $O/r02f_grid_lockstep.jsonl 2>&1
timeout 900 python tools/run_configs.py C1 C2 C3 C5 > $O/r02f_configs_gpu.jsonl 2>&1; echo "configs rc=$?"
for n in 8192 4096 1024; do timeout 200 python tools/dag_timeline.py $n $((n/16)) > $O/r02f_dag_timeline_n$n.txt 2>&1; done
timeout 100 ./build/bulk_probe 8192 4 > $O/r02f_bulk_probe.jsonl 2>&1
echo done
