#!/bin/bash
# A/B of search_kernel_fast variants (HNSWB200_LIB selects the build); correctness of the candidates first
mkdir -p gpurun_out
V=$PWD/hnsw_rs_b200/variants
for v in pipeq8 pipe; do
  (HNSWB200_LIB=$V/lib_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fast or heavy or search_synthetic or wide_rows or randomised") > gpurun_out/r2_pytest_d_$v.log 2>&1
  echo "pytest $v exit $?"; tail -2 gpurun_out/r2_pytest_d_$v.log
done
timeout 300 python bench.py --save-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_d_main.json 2> gpurun_out/r2_d_main.err
echo "bench main exit $?"
for v in norowpf pipe pipeq7 pipeq8 q7 q8; do
  HNSWB200_LIB=$V/lib_$v.so timeout 300 python bench.py --load-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_d_$v.json 2> gpurun_out/r2_d_$v.err
  echo "bench $v exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_d_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], 'value %.2fM'%(d['value']/1e6), 'ms %.4f'%d['ms_per_step'], 'alone %.4f'%r['kernel_ms_launched_alone'], 'frac', r['frac'], 'e2e %.2fM'%(d['e2e']['value']/1e6), 'spill', r.get('visited_spill_queries'), r['kernel'][-40:])
    except Exception as e: print(f, 'ERR', e)
PY
