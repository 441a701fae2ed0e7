#!/bin/bash
# Multi-GPU A/B of the peer exchange (gpurun --gpus N): pieces of the pipeline (TINYSPLAT_B200_PEER_CHUNKS),
# alternative builds (VARIANT_LIBS), correctness through tools/dp_check.py (also a ragged Gaussian count).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG="${TAG:-r2i}"; NG="${NG:-2}"; PORT=29700
mkdir -p gpurun_out
run() { timeout ${TMO:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((PORT++)) "$@"; }
if [ "${DPCHECK:-1}" = "1" ]; then
  run tools/dp_check.py --steps ${DPSTEPS:-20} > gpurun_out/${TAG}_dp_check_${NG}gpu.json 2> gpurun_out/${TAG}_dp_check_${NG}gpu.err
  tail -c 700 gpurun_out/${TAG}_dp_check_${NG}gpu.json; tail -3 gpurun_out/${TAG}_dp_check_${NG}gpu.err
  if [ "${RAGGED:-1}" = "1" ]; then
    run tools/dp_check.py --steps 5 --gaussians 777777 > gpurun_out/${TAG}_dp_check_${NG}gpu_ragged.json 2> gpurun_out/${TAG}_dp_check_${NG}gpu_ragged.err
    tail -c 400 gpurun_out/${TAG}_dp_check_${NG}gpu_ragged.json; tail -3 gpurun_out/${TAG}_dp_check_${NG}gpu_ragged.err
  fi
fi
for c in ${CHUNKS:-4 1 2}; do
  TINYSPLAT_B200_PEER_CHUNKS=$c run bench.py --gpus $NG --steps 30 --warmup 5 --grad-exchange peer --no-extras --sustained-s 0 \
     > gpurun_out/${TAG}_bench_${NG}gpu_peer_chunks$c.json 2> gpurun_out/${TAG}_bench_${NG}gpu_peer_chunks$c.err
  tail -c 200 gpurun_out/${TAG}_bench_${NG}gpu_peer_chunks$c.json; tail -2 gpurun_out/${TAG}_bench_${NG}gpu_peer_chunks$c.err
done
for lib in ${VARIANT_LIBS:-}; do
  name=$(basename $lib .so)
  export TINYSPLAT_B200_LIB=$PWD/$lib
  run tools/dp_check.py --steps 20 > gpurun_out/${TAG}_dp_check_${NG}gpu_$name.json 2> gpurun_out/${TAG}_dp_check_${NG}gpu_$name.err
  tail -c 400 gpurun_out/${TAG}_dp_check_${NG}gpu_$name.json
  for c in ${VARIANT_CHUNKS:-4 1}; do
    TINYSPLAT_B200_PEER_CHUNKS=$c run bench.py --gpus $NG --steps 30 --warmup 5 --grad-exchange peer --no-extras --sustained-s 0 \
       > gpurun_out/${TAG}_bench_${NG}gpu_${name}_chunks$c.json 2> gpurun_out/${TAG}_bench_${NG}gpu_${name}_chunks$c.err
    tail -c 200 gpurun_out/${TAG}_bench_${NG}gpu_${name}_chunks$c.json
  done
  unset TINYSPLAT_B200_LIB
done
if [ "${AUTO:-0}" = "1" ]; then
  NCCL_DEBUG=WARN run bench.py --gpus $NG --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_${NG}gpu_auto.json 2> gpurun_out/${TAG}_bench_${NG}gpu_auto.err
  tail -c 300 gpurun_out/${TAG}_bench_${NG}gpu_auto.json
fi
echo done
