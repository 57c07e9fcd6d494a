#!/bin/bash
# One gpurun call: GPU tests, micro-benchmarks, bench line, ncu launch list and ncu full captures.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [sections...]
#   sections: tests smoke optests engine ops mix batch bench parts launches ncu
set -u
TAG=${1:-rXX}; shift || true
SECTIONS=${*:-tests ops bench launches ncu}
O=gpurun_out/$TAG; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { echo BUILD FAILED; tail -20 $O/build.log; exit 1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/nvsmi.txt
for s in $SECTIONS; do
case $s in
tests)
  ( time timeout 1500 python -m pytest tests -m gpu -q -s ) > $O/pytest.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|error" $O/pytest.log | tail -5; grep -E "^(FAILED|ERROR)" $O/pytest.log | head -30 ;;
smoke)
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -3 $O/smoke.log ;;
optests)
  timeout 600 python -m pytest tests/test_unet_ops_gpu.py tests/test_gemm_gpu.py tests/test_unet_gpu.py tests/test_vae_gpu.py tests/test_mix_step_gpu.py -x -q -m "gpu and not slow" > $O/pytest_ops.log 2>&1; echo "op tests exit $?"; tail -3 $O/pytest_ops.log ;;
engine)
  timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -m gpu > $O/pytest_engine.log 2>&1; echo "engine tests exit $?"; tail -3 $O/pytest_engine.log ;;
mix)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/ubench_mix.cu -o tools/ubench_mix > $O/ubench_build.log 2>&1
  timeout 300 tools/ubench_mix > $O/ubench_mix.txt 2>&1; echo "ubench exit $?"; grep -c "mismatching_elements\": 0" $O/ubench_mix.txt; grep "rows\": 1984" $O/ubench_mix.txt ;;
batch)
  timeout 300 python tools/time_unet_batch.py > $O/unet_batch.txt 2>&1; cat $O/unet_batch.txt ;;
ops)
  timeout 300 python tools/bench_ops.py attn gemm > $O/bench_ops.txt 2>&1; cat $O/bench_ops.txt
  timeout 300 python tools/bench_mix.py > $O/bench_mix.txt 2>&1; cat $O/bench_mix.txt ;;
bench)
  timeout 1200 python bench.py --steps 2 --warmup 2 ${BENCH_ARGS:-} > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; cat $O/bench.json; tail -3 $O/bench.err ;;
parts)
  timeout 600 python tools/time_parts.py > $O/time_parts.txt 2>&1; head -12 $O/time_parts.txt ;;
launches)
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
     --log-file $O/unet_launches.csv python tools/profile_unet.py > $O/unet_launches.log 2>&1; echo "launches exit $?" ;;
traffic)
  timeout 900 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
     --clock-control none --csv --log-file $O/unet_traffic.csv python tools/profile_unet.py > $O/unet_traffic.log 2>&1; echo "traffic exit $?"
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
     -k regex:slerp_l2 -s 105 -c 2 --log-file $O/mix_traffic.csv python tools/bench_mix.py > $O/mix_traffic.log 2>&1; echo "mix traffic exit $?" ;;
small)
  timeout 300 python tools/bench_small.py > $O/bench_small.txt 2>&1; cat $O/bench_small.txt ;;
ncu_small)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gn_|ln_kernel|cfg_euler|scale_input|embed_inputs" -c 24 -o $O/small -f \
     env BENCH_ITERS=1 BENCH_WARM=0 python tools/bench_small.py > $O/ncu_small.log 2>&1; echo "ncu small exit $?" ;;
ncu_pair)
  for c in 1 2; do
    LB_GEMM_CLUSTER=$c timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o $O/ffin_cl$c -f \
       env BENCH_ITERS=1 BENCH_WARM=0 python tools/bench_ops.py gemm > $O/ncu_pair_$c.log 2>&1; echo "ncu pair $c exit $?"
  done ;;
gemmsplit)
  # which side bounds a GEMM: LB_GEMM_DEBUG=1 skips the TMA loads (MMA + epilogue only), =2 skips the MMA issue (TMA + epilogue only)
  : > $O/gemm_split.txt
  for d in 0 1 2; do echo "== LB_GEMM_DEBUG=$d" >> $O/gemm_split.txt; LB_GEMM_DEBUG=$d timeout 200 python tools/bench_ops.py gemm >> $O/gemm_split.txt 2>&1; done
  cat $O/gemm_split.txt ;;
ncu_gemm)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 13 -o $O/gemm -f \
     env BENCH_ITERS=1 BENCH_WARM=0 python tools/bench_ops.py gemm > $O/ncu_gemm.log 2>&1; echo "ncu gemm exit $?" ;;
ncu_attn)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -c 2 -o $O/attn -f \
     env BENCH_ITERS=1 BENCH_WARM=0 python tools/bench_ops.py attn > $O/ncu_attn.log 2>&1; echo "ncu attn exit $?" ;;
vaeprof)
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
     --log-file $O/vae_launches.csv python tools/profile_vae.py > $O/vae_launches.log 2>&1; echo "vae launches exit $?"; cat $O/vae_launches.log | tail -2
  LB_NO_GRAPH=0 timeout 300 python tools/profile_vae.py > $O/vae_time.txt 2>&1; cat $O/vae_time.txt ;;
configs)
  for c in 3 5 4; do
    timeout 900 python bench.py --config $c --steps ${CFG_STEPS:-2} --warmup ${CFG_WARM:-1} --no-cpu-baseline > $O/bench_config$c.json 2> $O/bench_config$c.err
    echo "config $c exit $?"; cut -c1-900 $O/bench_config$c.json; tail -2 $O/bench_config$c.err
  done ;;
attnpoly)
  : > $O/attn_poly.txt
  for v in 8 0 10 14; do echo "== LB_ATTN_POLY=$v" >> $O/attn_poly.txt; LB_ATTN_POLY=$v timeout 200 python tools/bench_ops.py attn >> $O/attn_poly.txt 2>&1; done
  cat $O/attn_poly.txt ;;
vaeconv)
  : > $O/bench_vae_conv.txt
  for c in "" "LB_GEMM_CLUSTER=1" "LB_GEMM_CLUSTER=2"; do env $c timeout 300 python tools/bench_ops.py vae >> $O/bench_vae_conv.txt 2>&1; done
  cat $O/bench_vae_conv.txt ;;
cluster2)
  LB_GEMM_CLUSTER=2 timeout 300 python tools/bench_ops.py gemm > $O/bench_ops_cluster2.txt 2>&1; cat $O/bench_ops_cluster2.txt ;;
dual)
  timeout 300 python tools/time_dual_stream.py > $O/dual_stream.txt 2>&1; cat $O/dual_stream.txt
  LB_GEMM_CORESIDENT=0 timeout 300 python tools/time_dual_stream.py > $O/dual_stream_nocr.txt 2>&1; cat $O/dual_stream_nocr.txt ;;
geglu)
  : > $O/geglu_ab.txt
  for e in "" "LB_GEGLU_TILE=256" "" "LB_GEGLU_TILE=256"; do env $e timeout 300 python tools/time_unet_batch.py 2 >> $O/geglu_ab.txt 2>&1; done
  cat $O/geglu_ab.txt ;;
ab)
  # same-box A/B of the UNet step (batch 2): default | no LN fold | no graph | GEGLU tile 256 | no PDL
  : > $O/ab.txt
  for e in "" "LB_NO_GRAPH=1" "LB_LN_FOLD=1" "LB_LN_FOLD=1 LB_NO_GRAPH=1" ${AB_EXTRA:-}; do
    env $e timeout 300 python tools/time_unet_batch.py 2 >> $O/ab.txt 2>&1
  done
  cat $O/ab.txt ;;
dual2)
  for e in "" "LB_NO_GRAPH=1"; do
    echo "== env: $e" >> $O/dual2.txt; env $e timeout 300 python tools/time_dual_stream.py >> $O/dual2.txt 2>&1
  done
  cat $O/dual2.txt ;;
lnfold)
  timeout 300 python tools/time_unet_batch.py > $O/unet_batch_lnfold.txt 2>&1; cat $O/unet_batch_lnfold.txt
  LB_NO_LN_FOLD=1 timeout 300 python tools/time_unet_batch.py > $O/unet_batch_nofold.txt 2>&1; cat $O/unet_batch_nofold.txt ;;
refarm)
  ( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref exit $?"; cat $O/bench_ref.json; tail -4 $O/bench_ref.err ;;
ncu)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -c 5 -o $O/attn -f \
     env BENCH_ITERS=1 BENCH_WARM=0 python tools/bench_ops.py attn > $O/ncu_attn.log 2>&1; echo "ncu attn exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 12 -o $O/gemm -f \
     env BENCH_ITERS=1 BENCH_WARM=0 python tools/bench_ops.py gemm > $O/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:slerp_l2 -s 105 -c 1 -o $O/mix -f \
     python tools/bench_mix.py > $O/ncu_mix.log 2>&1; echo "ncu mix exit $?" ;;
esac
done
ls -la $O
