O=gpurun_out
for cfg in "GPEDM_DAG_TRSM=4" "GPEDM_DAG_TRSM=6" "GPEDM_DAG_TRSM=8" "GPEDM_DAG_TRSM=12" "GPEDM_DAG_TRSM=999" "GPEDM_DAG_TRSM=0"; do
  env $cfg timeout -k 5 200 python tools/dag_sweep.py 512 1024 2048 3072 4096 6144 8192 12288; echo "sweep [$cfg] rc=$?"
done > $O/r02h_trsm_depth.jsonl 2>&1
python - <<'PY'
import json
rows={}
for l in open('gpurun_out/r02h_trsm_depth.jsonl'):
    if l.startswith('{'):
        r=json.loads(l); rows.setdefault(r['env'].get('GPEDM_DAG_TRSM'),{})[r['N']]=r['ms']
Ns=sorted(next(iter(rows.values())).keys())
print('depth '+' '.join('%8d'%n for n in Ns))
for k,v in rows.items(): print('%5s '%k+' '.join('%8.3f'%v.get(n,0) for n in Ns))
PY
for cfg in "GPEDM_DAG_TRSM=4" "GPEDM_DAG_TRSM=999" "GPEDM_DAG_TRSM=0"; do
  for rep in 1 2; do
  env $cfg timeout -k 5 120 python tools/mid_batch_bench.py 500 256 | sed "s/^{/{\"cfg\": \"$cfg\", /"
  env $cfg timeout -k 5 120 python tools/mid_batch_bench.py 2050 48 | sed "s/^{/{\"cfg\": \"$cfg\", /"
  done
done > $O/r02h_mid_depth.jsonl 2>&1
cut -c1-200 $O/r02h_mid_depth.jsonl
