#!/bin/bash
# BASELINE.json configs[3]: 1,048,576 envs over 8 GPUs (131,072 per GPU), plus the default 4096-envs/GPU line at 8 GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
NG=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $NG --steps 50 --warmup 5 > gpurun_out/bench_${NG}gpu.json 2> gpurun_out/bench_${NG}gpu.err
tail -3 gpurun_out/bench_${NG}gpu.err; cat gpurun_out/bench_${NG}gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $NG --envs 131072 --steps 6 --warmup 3 > gpurun_out/bench_${NG}gpu_1Menvs.json 2> gpurun_out/bench_${NG}gpu_1Menvs.err
tail -3 gpurun_out/bench_${NG}gpu_1Menvs.err; cat gpurun_out/bench_${NG}gpu_1Menvs.json
