#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x -k "fast or heavy or search_synthetic or wide_rows or randomised or c3 or c5") > gpurun_out/r2_pytest_h.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r2_pytest_h.log
timeout 600 python bench.py --config c3 --save-index /tmp/ix3 --steps 50 --warmup 5 > gpurun_out/r2_c3_n1.json 2> gpurun_out/r2_c3_n1.err
echo "c3 exit $?"; tail -2 gpurun_out/r2_c3_n1.err
HNSWB200_NO_FAST=1 timeout 600 python bench.py --config c3 --load-index /tmp/ix3 --ef 100 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_c3_n1_nofast.json 2> gpurun_out/r2_c3_n1_nofast.err
echo "c3 nofast exit $?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_c3_n1.json'))+['gpurun_out/r2_c3_n1_nofast.json']:
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], 'value %.2fM'%(d['value']/1e6), 'ms %.4f'%d['ms_per_step'], 'frac', r['frac'], 'e2e %.2fM'%(d['e2e']['value']/1e6), 'spill', r.get('visited_spill_queries'), 'ovf', r.get('visited_overflow_queries'), r['per_query'], r['kernel'][-60:])
    except Exception as e: print(f, 'ERR', e)
PY
