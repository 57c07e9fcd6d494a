#!/bin/bash
TAG=${1:-it}; O=gpurun_out/$TAG; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo BUILD FAILED; tail -20 $O/build.log; exit 1; }
timeout 600 python -m pytest tests/test_unet_ops_gpu.py tests/test_unet_gpu.py -x -q -m "gpu and not slow" > $O/pytest_ops.log 2>&1; echo "pytest exit $?"; tail -4 $O/pytest_ops.log
timeout 300 python tools/bench_ops.py attn > $O/bench_ops.txt 2>&1; cat $O/bench_ops.txt
