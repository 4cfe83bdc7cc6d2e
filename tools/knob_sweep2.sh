#!/bin/bash
for n in 4096 8192 16384; do
for cfg in "" "GPEDM_POTRF_GROUP_MIN=16" "GPEDM_FILL_TAIL=40" "GPEDM_POTRF_GROUP_MIN=16 GPEDM_FILL_TAIL=40" "GPEDM_POTRF_GROUP_MIN=12 GPEDM_FILL_TAIL=40" "GPEDM_POTRF_GROUP_MIN=16 GPEDM_FILL_TAIL=40 GPEDM_POTRF_GROUP=4"; do
  echo -n "N=$n [$cfg] "
  env $cfg python tools/dev_bench.py $n 2>&1 | grep '"bench"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_eval'], d['phases']['potrf'], d['phases']['trtri'])"
done; done
