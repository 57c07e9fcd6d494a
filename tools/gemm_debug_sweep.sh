for cl in 0 1; do for dbg in 0 1 2 3; do
  if [ $cl = 1 ]; then export LB_GEMM_NO_CLUSTER=1; else unset LB_GEMM_NO_CLUSTER; fi
  echo "== no_cluster=$cl debug=$dbg"; LB_GEMM_DEBUG=$dbg python tools/bench_ops.py gemm 2>&1 | grep -E "ff_out\"|conv1280|to_out"
done; done
