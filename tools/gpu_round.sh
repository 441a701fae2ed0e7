#!/bin/bash
# One consolidated GPU session (gpurun budget is tight): full -m gpu test-suite, A/B bench of the
# two blend-kernel generations, launch list + one ncu --set full capture of the grouped kernels.
# Everything lands in gpurun_out/; nothing here is a bench value when run under ncu.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 480 python -m pytest tests -q -m gpu ) > gpurun_out/tests_gpu.log 2>&1
tail -5 gpurun_out/tests_gpu.log
for m in 0 3; do
  TS_BLEND_MODE=$m timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline \
      > gpurun_out/bench_mode$m.json 2> gpurun_out/bench_mode$m.err
  tail -c 400 gpurun_out/bench_mode$m.json
done
TS_BLEND_MODE=3 timeout 200 ncu --set full --clock-control none --import-source on \
    -k regex:blend_.*group -c 2 -f -o gpurun_out/r1b_blend_group \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
TS_BLEND_MODE=3 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/r1b_launches_group.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_launches.log 2>&1
TS_BLEND_MODE=3 timeout 200 python bench.py > gpurun_out/bench_default_group.json 2> gpurun_out/bench_default_group.err
echo done
