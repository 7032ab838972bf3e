#!/bin/bash
# same-box A/B of an environment switch on the whole step: bash tools/gpu_ab.sh VAR A B [envs]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
VAR=$1; A=$2; B=$3; ENVS=${4:-32768}
for rep in 1 2; do
for V in $A $B; do
  env $VAR=$V timeout 600 python bench.py --envs $ENVS --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$VAR=$V', 'value %.0f' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'raycast_ms %.3f' % d['raycast_ms'])"
done
done
