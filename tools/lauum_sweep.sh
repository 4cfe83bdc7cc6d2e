#!/bin/bash
# Sweep of the Sigma^-1 piece schedule of the persistent factorisation kernel (GPEDM_DAG_LAUUM_* knobs):
# evaluation time per size, one JSON line per (chunk, tail, N).  Values must be bit-identical across knobs.
out=${1:-gpurun_out/lauum_sweep.log}
: > $out
for cfg in "0 1" "2 1" "4 1" "4 2" "8 1" "8 2" "16 2"; do
  set -- $cfg
  GPEDM_DAG_LAUUM_MAX_NB=1000 GPEDM_DAG_LAUUM_CHUNK=$1 GPEDM_DAG_LAUUM_TAIL=$2 timeout 300 python tools/dag_sweep.py 512 1024 2048 3072 4096 6144 8192 >> $out 2>&1
done
