#!/bin/bash
# A/B of the hand-off variants on the Cholesky chain (dag_factor.cuh: GPEDM_DAG_NO_STRIPS, GPEDM_DAG_TRSM), one B200:
# evaluation time per size with the results' bits against the committed sweep, chain timeline with the raw marks of one
# block column, parity subset, lockstep batch rate.
O=gpurun_out
SIZES="512 1024 2048 4096 8192"
for cfg in "" "GPEDM_DAG_TRSM=3" "GPEDM_DAG_TRSM=4" "GPEDM_DAG_TRSM=1" "GPEDM_DAG_TRSM=0"; do
  env $cfg timeout -k 5 150 python tools/dag_sweep.py $SIZES; echo "sweep [$cfg] rc=$?"
done > $O/r02h_strips_sweep.jsonl 2>&1
cut -c1-140 $O/r02h_strips_sweep.jsonl
python - <<'PY'
import json
ref = {}
for l in open('profiles/r02_size_sweep_v5.jsonl'):
    if l.startswith('{'):
        r = json.loads(l); ref[r['N']] = r
bad = 0
for l in open('gpurun_out/r02h_strips_sweep.jsonl'):
    if l.startswith('{'):
        r = json.loads(l); o = ref.get(r['N'])
        if o is not None and (o['nll'], o['g0']) != (r['nll'], r['g0']):
            bad += 1
            a, b = float.fromhex(o['nll']), float.fromhex(r['nll']); c, d = float.fromhex(o['g0']), float.fromhex(r['g0'])
            print('different bits', r['env'], r['N'], 'nll rel %.2e  g0 rel %.2e' % (abs(a - b) / abs(a), abs(c - d) / abs(c)))
print('bit check vs r02_size_sweep_v5: %d differences' % bad)
PY
DAG_TL_DUMP=10 timeout -k 5 120 python tools/dag_timeline.py 4096 256 > $O/r02h_dag_timeline_n4086_strips.txt 2>&1; echo "timeline rc=$?"; grep "chain step:" $O/r02h_dag_timeline_n4086_strips.txt
timeout -k 5 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_midbatch.py -m gpu -x -q -k "likegrad or K1 or mid or lockstep or sequential" 2>&1 | tail -4
for cfg in "" "GPEDM_DAG_TRSM=0"; do
  env $cfg timeout -k 5 120 python tools/mid_batch_bench.py 1000 128 | sed "s/^{/{\"cfg\": \"$cfg\", /"
  env $cfg timeout -k 5 120 python tools/mid_batch_bench.py 3000 24 | sed "s/^{/{\"cfg\": \"$cfg\", /"
done > $O/r02h_mid_batch_strips.jsonl 2>&1
cut -c1-260 $O/r02h_mid_batch_strips.jsonl
echo done
