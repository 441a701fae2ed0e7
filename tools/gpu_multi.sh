#!/bin/bash
# Multi-GPU session (gpurun --gpus N; charged N x): correctness of the gradient-exchange strategies
# (tools/dp_check.py: all-reduce vs packed vs peer) and bench lines per strategy.  Everything runs
# under `timeout` so that a hung collective cannot burn the GPU budget.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG="${TAG:-r2c}"
NG="${NG:-2}"
STRATEGIES="${STRATEGIES:-auto allreduce packed peer}"
PORT=29600
mkdir -p gpurun_out
run() { timeout ${TMO:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((PORT++)) "$@"; }
if [ "${DPCHECK:-1}" = "1" ]; then
  run tools/dp_check.py --steps 20 > gpurun_out/${TAG}_dp_check_${NG}gpu.json 2> gpurun_out/${TAG}_dp_check_${NG}gpu.err
  tail -c 1500 gpurun_out/${TAG}_dp_check_${NG}gpu.json; tail -5 gpurun_out/${TAG}_dp_check_${NG}gpu.err
fi
for s in $STRATEGIES; do
  NCCL_DEBUG=${NCCL_DEBUG_LEVEL:-WARN} run bench.py --gpus $NG --steps 30 --warmup 5 --grad-exchange $s \
      > gpurun_out/${TAG}_bench_${NG}gpu_$s.json 2> gpurun_out/${TAG}_bench_${NG}gpu_$s.err
  tail -c 300 gpurun_out/${TAG}_bench_${NG}gpu_$s.json; tail -3 gpurun_out/${TAG}_bench_${NG}gpu_$s.err
done
for w in ${EXTRA_WORKLOADS:-}; do
  run bench.py --gpus $NG --steps 20 --warmup 5 --workload $w > gpurun_out/${TAG}_bench_${NG}gpu_$w.json 2> gpurun_out/${TAG}_bench_${NG}gpu_$w.err
done
echo done
