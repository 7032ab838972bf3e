cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python bench.py --envs 65536 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c3_c.json 2> gpurun_out/bench_c3_c.err; tail -2 gpurun_out/bench_c3_c.err
timeout 600 python bench.py --envs 131072 --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_c4share_c.json 2> gpurun_out/bench_c4share_c.err; tail -2 gpurun_out/bench_c4share_c.err
