#!/bin/bash
# One consolidated GPU session (the gpurun budget is tight): the -m gpu test-suite, the default
# bench line, A/B bench lines of experimental builds (VARIANT_LIBS: paths of alternative .so files,
# selected through TINYSPLAT_B200_LIB) and of environment switches (ENV_VARIANTS), the ncu launch
# list of the bench command, a torch.profiler table of the drop-in pipeline.
# Everything lands in gpurun_out/; nothing run under ncu is a bench value.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG="${TAG:-r2a}"
VARIANT_LIBS="${VARIANT_LIBS:-}"
ENV_VARIANTS="${ENV_VARIANTS:-}"     # e.g. "TINYSPLAT_B200_SH_BWD_PRIORITY=1 TS_BLEND_MODE=warp"
TESTS="${TESTS:-1}"
NCU_LIST="${NCU_LIST:-1}"
TORCH_PROFILE="${TORCH_PROFILE:-}"   # e.g. "reference fused"
AB="--steps 30 --warmup 5 --no-cpu-baseline --no-extras --sustained-s 0"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
if [ "$TESTS" = "1" ]; then
  ( time timeout 900 python -m pytest tests -q -m gpu -s ) > gpurun_out/${TAG}_tests_gpu.log 2>&1
  tail -5 gpurun_out/${TAG}_tests_gpu.log
fi
timeout 400 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
tail -c 400 gpurun_out/${TAG}_bench_default.json; tail -3 gpurun_out/${TAG}_bench_default.err
timeout 120 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference_arm.json 2>/dev/null
for lib in $VARIANT_LIBS; do
  name=$(basename $lib .so)
  TINYSPLAT_B200_LIB=$PWD/$lib timeout 120 python bench.py $AB \
      > gpurun_out/${TAG}_bench_$name.json 2> gpurun_out/${TAG}_bench_$name.err
done
for ev in $ENV_VARIANTS; do
  name=$(echo $ev | tr '=' '_')
  env $ev timeout 120 python bench.py $AB \
      > gpurun_out/${TAG}_bench_env_$name.json 2> gpurun_out/${TAG}_bench_env_$name.err
done
timeout 120 python bench.py $AB > gpurun_out/${TAG}_bench_ab_default.json 2>/dev/null
if [ "$NCU_LIST" = "1" ]; then
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
      --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --sustained-s 0 \
      > gpurun_out/${TAG}_ncu_launches.log 2>&1
fi
if [ -n "${NCU_FULL:-}" ]; then   # e.g. NCU_FULL="blend_fwd_pair|blend_bwd_group": one --set full capture of the first launches
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$NCU_FULL" -s ${NCU_SKIP:-8} -c ${NCU_COUNT:-2} \
      -f -o gpurun_out/${TAG}_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --sustained-s 0 \
      > gpurun_out/${TAG}_ncu_full.log 2>&1
  ncu -i gpurun_out/${TAG}_full.ncu-rep --page details --csv > gpurun_out/${TAG}_full_details.csv 2>/dev/null
fi
for pipe in $TORCH_PROFILE; do
  timeout 120 python tools/torch_profile.py synthetic_1M_1080p $pipe > gpurun_out/${TAG}_torch_profile_$pipe.txt 2>&1
done
echo done
