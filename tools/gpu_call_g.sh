#!/bin/bash
# correctness of the trimmed kernel + spill pool, then C3 on one GPU with 7 / 6 / 5 / 8 blocks per SM for the ef <= 128 variant, then C2 bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x -k "fast or heavy or search_synthetic or wide_rows or randomised or kernel_variants or c3 or c5") > gpurun_out/r2_pytest_g.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r2_pytest_g.log
timeout 600 python bench.py --config c3 --save-index /tmp/ix3 --steps 50 --warmup 5 > gpurun_out/r2_c3_n1.json 2> gpurun_out/r2_c3_n1.err
echo "c3 exit $?"; tail -2 gpurun_out/r2_c3_n1.err
for v in k4_6 k4_5 k4_8; do
  HNSWB200_LIB=$PWD/hnsw_rs_b200/variants/lib_$v.so timeout 600 python bench.py --config c3 --load-index /tmp/ix3 --ef 100 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_c3_n1_$v.json 2> gpurun_out/r2_c3_n1_$v.err
  echo "c3 $v exit $?"
done
timeout 400 python bench.py --no-cpu-baseline --ef 57 > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err
echo "bench c2 exit $?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_c3_n1*.json'))+['gpurun_out/r2_bench_g.json']:
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], 'value %.2fM'%(d['value']/1e6), 'ms %.4f'%d['ms_per_step'], 'frac', r['frac'], 'e2e %.2fM'%(d['e2e']['value']/1e6), 'spill', r.get('visited_spill_queries'), 'ovf', r.get('visited_overflow_queries'), r['kernel'][-44:])
    except Exception as e: print(f, 'ERR', e)
PY
