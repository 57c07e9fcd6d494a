#!/bin/bash
# mix-kernel iteration: A/B micro-benchmark + parity tests + library microbench
TAG=${1:-mix}; O=gpurun_out/$TAG; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo BUILD FAILED; tail -20 $O/build.log; exit 1; }
timeout 300 tools/ubench_mix > $O/ubench_mix.txt 2>&1; echo "ubench exit $?"; cat $O/ubench_mix.txt
timeout 600 python -m pytest tests/test_mix_step_gpu.py -x -q > $O/pytest_mix.log 2>&1; echo "pytest exit $?"; tail -5 $O/pytest_mix.log
timeout 300 python tools/bench_mix.py > $O/bench_mix.txt 2>&1; cat $O/bench_mix.txt
