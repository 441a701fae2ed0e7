#!/bin/bash
# One consolidated GPU session (the gpurun budget is tight): the -m gpu test-suite, the default
# bench line, A/B bench lines of experimental builds (VARIANT_LIBS: paths of alternative .so files,
# selected through TINYSPLAT_B200_LIB) and of environment switches (ENV_VARIANTS), the ncu launch
# list of the bench command, extra workloads.
# Everything lands in gpurun_out/; nothing run under ncu is a bench value.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG="${TAG:-r1i}"
VARIANT_LIBS="${VARIANT_LIBS:-}"
ENV_VARIANTS="${ENV_VARIANTS:-}"     # e.g. "TINYSPLAT_B200_SH_BWD_PRIORITY=1 TS_BLEND_MODE=warp"
EXTRA_WORKLOADS="${EXTRA_WORKLOADS:-}"
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -q -m gpu ) > gpurun_out/${TAG}_tests_gpu.log 2>&1
tail -4 gpurun_out/${TAG}_tests_gpu.log
timeout 200 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
tail -c 300 gpurun_out/${TAG}_bench_default.json
for lib in $VARIANT_LIBS; do
  name=$(basename $lib .so)
  TINYSPLAT_B200_LIB=$PWD/$lib timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline \
      > gpurun_out/${TAG}_bench_$name.json 2> gpurun_out/${TAG}_bench_$name.err
done
for ev in $ENV_VARIANTS; do
  name=$(echo $ev | tr '=' '_')
  env $ev timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline \
      > gpurun_out/${TAG}_bench_env_$name.json 2> gpurun_out/${TAG}_bench_env_$name.err
done
timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_ab_default.json 2>/dev/null
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_launches.log 2>&1
for w in $EXTRA_WORKLOADS; do
  timeout 150 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline \
      > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
done
echo done
