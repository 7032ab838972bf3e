#!/bin/bash
# compute-sanitizer (memcheck, then racecheck) over a small slice of the GPU parity tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
SEL='golden_get_depths or golden_rock_detection or golden_task_step or fused_step or reset_targets_matches'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/sanitize_memcheck.txt 2>&1
echo "memcheck rc=$?"; tail -8 gpurun_out/sanitize_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests -m gpu -q -x -k "golden_get_depths or fused_step" > gpurun_out/sanitize_racecheck.txt 2>&1
echo "racecheck rc=$?"; tail -12 gpurun_out/sanitize_racecheck.txt
