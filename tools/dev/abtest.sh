timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --save-index /tmp/ix > gpurun_out/b.log 2>&1
for v in A B C D E; do
  if [ $v = A ]; then unset HNSWB200_LIB; else export HNSWB200_LIB=$PWD/hnsw_rs_b200/variants/lib_$v.so; fi
  for nq in 10000 100000; do
    echo "variant $v nq $nq"; timeout 300 python tools/dev/exp_search.py --load /tmp/ix --nq $nq --efs 64 --oracle-sample 0 2>&1 | grep "ef="
  done
done
