#!/bin/bash
# fast-kernel tests on the new code, then A/B of the variants on the headline workload (device-resident numbers only)
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fast or heavy or kernel_variants or search_synthetic or wide_rows or randomised or glove_fixture") > gpurun_out/r2_pytest_c.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/r2_pytest_c.log
timeout 300 python bench.py --save-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_c_main.json 2> gpurun_out/r2_c_main.err
echo "bench main exit $?"
for v in nospec oldmerge nopf; do
  HNSWB200_LIB=$PWD/hnsw_rs_b200/variants/lib_$v.so timeout 300 python bench.py --load-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_c_$v.json 2> gpurun_out/r2_c_$v.err
  echo "bench $v exit $?"
done
for b in 6 5; do
  HNSWB200_SEARCH_BLOCKS_PER_SM=$b timeout 300 python bench.py --load-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_c_blocks$b.json 2> gpurun_out/r2_c_blocks$b.err
  echo "bench blocks $b exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_c_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], 'value %.2fM'%(d['value']/1e6), 'ms %.4f'%d['ms_per_step'], 'alone %.4f'%r['kernel_ms_launched_alone'], 'frac', r['frac'], 'e2e %.2fM'%(d['e2e']['value']/1e6), 'spill', r.get('visited_spill_queries'))
    except Exception as e: print(f, 'ERR', e)
PY
