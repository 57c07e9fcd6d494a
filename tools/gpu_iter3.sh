#!/bin/bash
TAG=${1:-it}; O=gpurun_out/$TAG; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo BUILD FAILED; tail -20 $O/build.log; exit 1; }
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -m "gpu" > $O/pytest_engine.log 2>&1; echo "pytest exit $?"; tail -6 $O/pytest_engine.log
timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; cat $O/bench.json; tail -3 $O/bench.err
