#!/bin/bash
# One multi-GPU gpurun call: the sharded-engine exactness tests and a short sharded bench.
# usage (under gpurun --gpus N): bash tools/gpu_multi.sh <tag> <N> [config] [steps] [warmup]
set -u
TAG=${1:-rXX}; N=${2:-2}; CFG=${3:-2}; STEPS=${4:-2}; WARM=${5:-2}
O=gpurun_out/$TAG; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo BUILD FAILED; tail -20 $O/build.log; exit 1; }
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > $O/nvsmi.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  ( time timeout 900 python -m pytest tests/test_sharded_engine_gpu.py -q -s -m gpu ) > $O/pytest_sharded.log 2>&1
  echo "sharded tests exit $?"; grep -E "passed|failed|skipped|shard stats|world" $O/pytest_sharded.log | tail -8
fi
for n in $N; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 \
     bench.py --gpus $n --config $CFG --steps $STEPS --warmup $WARM > $O/bench_n$n.json 2> $O/bench_n$n.err
  echo "bench N=$n exit $?"; cat $O/bench_n$n.json | cut -c1-1500; tail -3 $O/bench_n$n.err
done
