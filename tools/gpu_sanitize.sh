#!/bin/bash
# compute-sanitizer (memcheck, then racecheck) over a slice of the GPU parity tests: ray-cast (shadow + tiled + compact layer), rock,
# fused step, device reset path, tensor-core policy kernels, hooks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
TAG=${1:-r2}
SEL='rock_kernel_candidate_counts or degenerate or golden_get_depths or golden_rock_detection or golden_task_step or fused_step or reset_targets_matches or compact_layer or edge_cases or pre_physics_step_device or hooks_agree'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/sanitize_memcheck_$TAG.txt 2>&1
echo "memcheck rc=$?"; tail -8 gpurun_out/sanitize_memcheck_$TAG.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_zz_gpu_policy.py -m gpu -q -x -k "golden_reference_outputs or pair_launch or strided_rows" > gpurun_out/sanitize_memcheck_policy_$TAG.txt 2>&1
echo "memcheck policy rc=$?"; tail -8 gpurun_out/sanitize_memcheck_policy_$TAG.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests -m gpu -q -x -k "golden_get_depths or fused_step or compact_layer" > gpurun_out/sanitize_racecheck_$TAG.txt 2>&1
echo "racecheck rc=$?"; tail -12 gpurun_out/sanitize_racecheck_$TAG.txt
