#!/bin/bash
# Final evidence of round 2 on the last tree (one B200): full GPU test suite, smoke, bench line, ncu launch list of the
# bench command, evaluation time vs size, configs C1 / C2.
O=gpurun_out
timeout -k 5 600 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -10 | tee $O/r02j_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 > $O/r02j_bench.json 2> $O/r02j_bench.err; echo "bench rc=$?"; cut -c1-330 $O/r02j_bench.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02j_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-grid --no-cpu-baseline > $O/r02j_ncu_bench.log 2>&1; echo "ncu list rc=$?"
python tools/summarize_launches.py $O/r02j_launches_bench.csv > $O/r02j_launch_summary_bench.txt 2>&1; head -12 $O/r02j_launch_summary_bench.txt
timeout 300 python tools/dag_sweep.py 512 1024 2048 3072 4096 6144 8192 12288 16384 24576 32768 > $O/r02j_size_sweep.jsonl 2>&1; echo "sweep rc=$?"; cut -c1-100 $O/r02j_size_sweep.jsonl
timeout 200 python tools/run_configs.py C1 C2 > $O/r02j_configs.jsonl 2>&1; echo "configs rc=$?"; cut -c1-300 $O/r02j_configs.jsonl
echo done
