#!/bin/bash
# one-shot check of a timing-only change to the chain hand-off: the bits of the results against the committed sweep of the
# final tree (must be identical), evaluation times, and the parity tests that exercise the persistent kernel most
timeout -k 5 40 python tools/dag_sweep.py 1024 2048 4096 8192 > gpurun_out/r02m_sweep.jsonl 2>&1; echo "sweep rc=$?"
python - <<'PY'
import json
ref = {json.loads(l)['N']: json.loads(l) for l in open('profiles/r02_size_sweep_v6.jsonl') if l.startswith('{')}
bad = 0
for l in open('gpurun_out/r02m_sweep.jsonl'):
    if l.startswith('{'):
        r = json.loads(l); o = ref[r['N']]
        same = (o['nll'], o['g0']) == (r['nll'], r['g0']); bad += not same
        print(r['N'], o['ms'], '->', r['ms'], 'bit-identical' if same else 'DIFFERENT')
print('differences:', bad)
PY
timeout -k 5 60 python -m pytest tests/test_gpu_parity.py tests/test_gpu_midbatch.py -m gpu -x -q -k "likegrad or K1 or matches_tiled" 2>&1 | tail -2
