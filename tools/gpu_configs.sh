#!/bin/bash
# BASELINE.json configs[2..4] on one GPU: 65,536 envs, 131,072 envs (one rank's share of the 1M-env job), C5 sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
TAG=${1:-cfg}
timeout 600 python bench.py --envs 65536 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err
tail -3 gpurun_out/bench_c3_$TAG.err; cat gpurun_out/bench_c3_$TAG.json
timeout 600 python bench.py --envs 131072 --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_c4share_$TAG.json 2> gpurun_out/bench_c4share_$TAG.err
tail -3 gpurun_out/bench_c4share_$TAG.err; cat gpurun_out/bench_c4share_$TAG.json
timeout 1200 python tools/sweep_c5.py --envs 4096 > gpurun_out/sweep_c5_$TAG.txt 2> gpurun_out/sweep_c5_$TAG.err
tail -5 gpurun_out/sweep_c5_$TAG.err; cat gpurun_out/sweep_c5_$TAG.txt
