cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
TAG=$1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hm_shadow -s 3 -c 1 -f -o gpurun_out/prof_$TAG python bench.py --envs 4096 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
