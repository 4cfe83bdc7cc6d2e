timeout 300 python tools/dag_sweep.py 512 1024 2048 3072 4096 6144 8192 > gpurun_out/r02_sweep_after3.log 2>&1
timeout 200 python tools/dag_timeline.py 1024 50 > gpurun_out/r02_dag_tl_marks_1024f.txt 2>&1
timeout 200 python tools/dag_timeline.py 4096 100 > gpurun_out/r02_dag_tl_marks_4096f.txt 2>&1
python - <<'PY'
import json
b={json.loads(l)['N']:json.loads(l) for l in open('tools/_before.log')}
for l in open('gpurun_out/r02_sweep_after3.log'):
    if l.startswith('{'):
        r=json.loads(l); o=b[r['N']]
        print(r['N'], o['ms'], '->', r['ms'], 'bit-identical' if (o['nll'],o['g0'])==(r['nll'],r['g0']) else 'DIFFERENT')
PY
grep -A12 "chain step breakdown" gpurun_out/r02_dag_tl_marks_1024f.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "likegrad or K1 or predict_new or sequential" 2>&1 | tail -3
