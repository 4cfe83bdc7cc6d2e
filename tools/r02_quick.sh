#!/bin/bash
# quick A/B after a change to the persistent kernel: evaluation time per size (default / product with D_c), bench line, parity subset
O=gpurun_out
for cfg in "" "GPEDM_DAG_TRSM=0"; do
  env $cfg timeout -k 5 150 python tools/dag_sweep.py 1024 2048 4096 8192 | cut -c1-110; echo "sweep [$cfg] rc=$?"
done
timeout 200 python bench.py --steps 10 --warmup 3 --no-grid --no-cpu-baseline 2>/dev/null | cut -c1-230
timeout -k 5 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_midbatch.py -m gpu -x -q -k "likegrad or K1 or mid or lockstep or sequential" 2>&1 | tail -2
