#!/bin/bash
# strip-split thresholds of the chain kernels (factor in halves of the SM count: 4 = tiles*4 <= 2*SMs)
for n in 8192 4096; do
for cfg in "" "GPEDM_SPLIT_UPDATE2=2" "GPEDM_SPLIT_UPDATE2=3" "GPEDM_SPLIT_PANEL2=2" "GPEDM_SPLIT_PANEL2=3" "GPEDM_SPLIT_UPDATE2=2 GPEDM_SPLIT_PANEL2=2" "GPEDM_SPLIT_UPDATE2=6 GPEDM_SPLIT_PANEL2=6"; do
  echo -n "N=$n [$cfg] "
  env $cfg python tools/dev_bench.py $n 2>&1 | grep '"bench"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_eval'], d['phases']['potrf'], d['phases']['trtri'])"
done; done
