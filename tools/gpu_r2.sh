#!/bin/bash
# round-2 GPU call: diagnostic, parity tests, bench lines (4096 and 131072 envs); optional ncu passes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
TAG=${1:-run}
KERN=${3:-hm_shadow}
timeout 600 python tools/shadow_diag.py 4096 > gpurun_out/diag_$TAG.txt 2>&1
tail -32 gpurun_out/diag_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu_$TAG.txt 2>&1
tail -8 gpurun_out/pytest_gpu_$TAG.txt
timeout 600 python bench.py --envs 4096 --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_4096_$TAG.json 2> gpurun_out/bench_4096_$TAG.err
tail -3 gpurun_out/bench_4096_$TAG.err; cut -c1-400 gpurun_out/bench_4096_$TAG.json
timeout 600 python bench.py --envs 131072 --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_131072_$TAG.json 2> gpurun_out/bench_131072_$TAG.err
tail -3 gpurun_out/bench_131072_$TAG.err; cut -c1-400 gpurun_out/bench_131072_$TAG.json
if [ "$2" == "prof" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --envs 4096 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KERN -s 3 -c 2 -f -o gpurun_out/prof_$TAG \
    python bench.py --envs 4096 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
fi
