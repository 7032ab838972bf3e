#!/bin/bash
# round-2 GPU call: [diag] parity tests, smoke, bench lines; optional ncu passes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
TAG=${1:-run}
MODE=${2:-all}
KERN=${3:-hm_shadow}
if [[ "$MODE" == *diag* || "$MODE" == all ]]; then
timeout 600 python tools/shadow_diag.py 4096 > gpurun_out/diag_$TAG.txt 2>&1
tail -34 gpurun_out/diag_$TAG.txt
fi
if [[ "$MODE" == *test* || "$MODE" == all ]]; then
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu_$TAG.txt 2>&1
tail -25 gpurun_out/pytest_gpu_$TAG.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.txt 2>&1
tail -3 gpurun_out/smoke_$TAG.txt
fi
if [[ "$MODE" == *bench* || "$MODE" == all ]]; then
timeout 900 python bench.py > gpurun_out/bench_default_$TAG.json 2> gpurun_out/bench_default_$TAG.err
tail -3 gpurun_out/bench_default_$TAG.err; cat gpurun_out/bench_default_$TAG.json
timeout 600 python bench.py --envs 4096 --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_4096_$TAG.json 2> gpurun_out/bench_4096_$TAG.err
tail -3 gpurun_out/bench_4096_$TAG.err; cut -c1-300 gpurun_out/bench_4096_$TAG.json
fi
if [[ "$MODE" == *ref* ]]; then
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
tail -3 gpurun_out/bench_ref_$TAG.err; cat gpurun_out/bench_ref_$TAG.json
fi
if [[ "$MODE" == *prof* ]]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --envs 4096 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KERN -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --envs 4096 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
fi
