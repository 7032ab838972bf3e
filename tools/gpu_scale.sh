#!/bin/bash
# the driver's scaling run: default workload at N GPUs, our arm (+ the reference arm when $2 = ref)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
NG=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/bench_${NG}gpu.json 2> gpurun_out/bench_${NG}gpu.err
tail -3 gpurun_out/bench_${NG}gpu.err; cat gpurun_out/bench_${NG}gpu.json
if [ "$2" == "ref" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $NG --steps 3 --warmup 1 > gpurun_out/bench_${NG}gpu_ref.json 2> gpurun_out/bench_${NG}gpu_ref.err
cat gpurun_out/bench_${NG}gpu_ref.json | cut -c1-300
fi
