#!/bin/bash
TAG=${1:-it}; O=gpurun_out/$TAG; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo BUILD FAILED; tail -20 $O/build.log; exit 1; }
timeout 600 python -m pytest tests/test_unet_ops_gpu.py tests/test_unet_gpu.py tests/test_vae_gpu.py -x -q -m "gpu and not slow" > $O/pytest_ops.log 2>&1; echo "pytest exit $?"; tail -4 $O/pytest_ops.log
timeout 300 python tools/time_unet_batch.py > $O/unet_batch.txt 2>&1; cat $O/unet_batch.txt
timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; cat $O/bench.json; tail -3 $O/bench.err
