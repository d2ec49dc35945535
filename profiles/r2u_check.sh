#!/bin/bash
# profiles/r2u_check.sh -- host-packed upload with a character share beside it: tests, then e2e per packed share
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hostpack.py tests/test_gpu_parity.py -m gpu -x -q -k "hostpack or chunk_streamed or golden" > gpurun_out/r2u_pytest.txt 2>&1; tail -3 gpurun_out/r2u_pytest.txt
for f in 0.4 0.5 0.6 0.7; do echo "== GPUMOTIF_PACK_FRAC=$f"; GPUMOTIF_PACK_FRAC=$f bash profiles/quickbench.sh frac$f 1024 trna ire; done
echo "== default"; bash profiles/quickbench.sh fracd 1024 trna ire score.1 qu+tr pk1
python - <<'PY'
import json
for t in ['0.4','0.5','0.6','0.7','d']:
    for d in ['trna','ire']:
        j=json.load(open(f'gpurun_out/qb_frac{t}_{d}.json')); e=j['e2e']
        print(t, d, 'e2e ms', round(e['ms_per_step'],2), e['upload'][:8], {k:round(v['ms_per_step'],2) for k,v in e['other_uploads'].items()}, 'h2d', e['h2d_bytes_per_step'], e['phases_ms_last_step']['h2d_ms'])
PY
