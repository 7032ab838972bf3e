#!/bin/bash
# first GPU call: semantics probe + parity tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
python tests/gpu_probe.py > gpurun_out/probe.txt 2>&1
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.txt 2>&1
tail -30 gpurun_out/probe.txt
tail -40 gpurun_out/pytest_gpu.txt
