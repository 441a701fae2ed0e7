#!/bin/bash
# One consolidated GPU session (gpurun budget is tight): full -m gpu test-suite, A/B bench of the
# blend-kernel variants (MODES, see ts_set_blend_mode), launch list + one ncu --set full capture of
# the blend kernels of NCU_MODE.  Everything lands in gpurun_out/; nothing run under ncu is a bench value.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
MODES="${MODES:-0 2 4}"
NCU_MODE="${NCU_MODE:-4}"
TAG="${TAG:-r1c}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 480 python -m pytest tests -q -m gpu ) > gpurun_out/${TAG}_tests_gpu.log 2>&1
tail -5 gpurun_out/${TAG}_tests_gpu.log
for m in $MODES; do
  TS_BLEND_MODE=$m timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline \
      > gpurun_out/${TAG}_bench_mode$m.json 2> gpurun_out/${TAG}_bench_mode$m.err
  tail -c 300 gpurun_out/${TAG}_bench_mode$m.json
done
TS_BLEND_MODE=$NCU_MODE timeout 200 ncu --set full --clock-control none --import-source on \
    -k regex:blend_ -c 2 -f -o gpurun_out/${TAG}_blend_mode$NCU_MODE \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
TS_BLEND_MODE=$NCU_MODE timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/${TAG}_launches_mode$NCU_MODE.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo done
