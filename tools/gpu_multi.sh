#!/bin/bash
# multi-GPU configurations on N GPUs of one box: usage gpu_multi.sh N "c3 c4 c5" [extra bench args for c5]
# (N = 1 runs them in-process).  Writes gpurun_out/r2_<cfg>_n<N>.json; with N = 2 also the 2-GPU pytest log.
N=$1; CFGS=$2; shift 2
mkdir -p gpurun_out
run() {
  cfg=$1; shift
  if [ "$N" = "1" ]; then
    timeout 1500 python bench.py --config $cfg --gpus 1 "$@" > gpurun_out/r2_${cfg}_n1.json 2> gpurun_out/r2_${cfg}_n1.err
  else
    timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --config $cfg --gpus $N "$@" > gpurun_out/r2_${cfg}_n$N.json 2> gpurun_out/r2_${cfg}_n$N.err
  fi
  echo "$cfg N=$N exit $?"; tail -c 600 gpurun_out/r2_${cfg}_n$N.json; echo; tail -3 gpurun_out/r2_${cfg}_n$N.err
}
if [ "$N" = "2" ]; then
  (timeout 900 python -m pytest tests/test_sharded_gpu.py -m gpu -q -x -rs) > gpurun_out/r2_pytest_2gpu.log 2>&1
  echo "2-GPU pytest exit $?"; tail -5 gpurun_out/r2_pytest_2gpu.log
fi
for c in $CFGS; do
  if [ "$c" = "c5" ]; then run c5 "$@"; else run $c; fi
done
