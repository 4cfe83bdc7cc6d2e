#!/bin/bash
# Sweep the scheduling knobs of the Cholesky (env vars read at library load) at N=8182.
for cfg in "" "GPEDM_POTRF_GROUP=1" "GPEDM_POTRF_GROUP=3" "GPEDM_POTRF_GROUP=4" "GPEDM_POTRF_GROUP_MIN=16" "GPEDM_POTRF_GROUP_MIN=32" \
           "GPEDM_FILL_TAIL=24" "GPEDM_FILL_TAIL=40" "GPEDM_FILL_TAIL=48" "GPEDM_FILL_TAIL=64" "GPEDM_NO_FILL=1" "GPEDM_NO_SPLIT=1" "GPEDM_NO_HALF_TILES=1"; do
  echo -n "[$cfg] "
  env $cfg python tools/dev_bench.py ${1:-8192} 2>&1 | grep '"bench"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_eval'], d['phases']['potrf'], d['phases']['trtri'])"
done
