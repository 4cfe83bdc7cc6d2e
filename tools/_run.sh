timeout 60 python tools/dag_sweep.py 256 384 512 > gpurun_out/r02_sweep_chain_small.log 2>&1; echo "small rc=$?"; cut -c1-150 gpurun_out/r02_sweep_chain_small.log
timeout 300 python tools/dag_sweep.py 512 1024 2048 3072 4096 6144 8192 > gpurun_out/r02_sweep_after4.log 2>&1; echo "sweep rc=$?"
GPEDM_DAG_CHAIN_MAX_NB=100 timeout 100 python tools/dag_sweep.py 6144 8192 2>&1 | cut -c1-120
python - <<'PY'
import json
b={json.loads(l)['N']:json.loads(l) for l in open('tools/_before.log')}
for l in open('gpurun_out/r02_sweep_after4.log'):
    if l.startswith('{'):
        r=json.loads(l); o=b[r['N']]
        print(r['N'], o['ms'], '->', r['ms'], 'bit-identical' if (o['nll'],o['g0'])==(r['nll'],r['g0']) else 'DIFFERENT')
PY
timeout 100 python tools/dag_timeline.py 1024 50 > gpurun_out/r02_dag_tl_1024_chain.txt 2>&1; tail -4 gpurun_out/r02_dag_tl_1024_chain.txt
timeout 100 python tools/dag_timeline.py 4096 100 > gpurun_out/r02_dag_tl_4096_chain.txt 2>&1; tail -2 gpurun_out/r02_dag_tl_4096_chain.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "likegrad or K1 or predict_new or sequential" 2>&1 | tail -3
