#!/bin/bash
# tools/run_multigpu.sh N TAG [configs...] -- bench.py on N GPUs of one box for the given configs (default: metric c4 c5),
# records under gpurun_out/<TAG>_bench_<config>_<N>gpu.json (development tool; the driver's own invocation is the
# plain torch.distributed.run line in bench.py's docstring).
N=$1; TAG=$2; shift 2
CFGS=${@:-metric c4 c5}
mkdir -p gpurun_out
for cfg in $CFGS; do
  steps=5; [ "$cfg" != metric ] && steps=3
  out=gpurun_out/${TAG}_bench_${cfg}_${N}gpu
  if [ "$N" = 1 ]; then
    timeout 300 python bench.py --config $cfg --steps $steps --warmup 3 $BENCH_EXTRA > $out.json 2> $out.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
      bench.py --gpus $N --config $cfg --steps $steps --warmup 3 $BENCH_EXTRA > $out.json 2> $out.err
  fi
  echo "== $cfg N=$N rc=$?"; head -c 400 $out.json; echo; tail -2 $out.err
done
