for c in 0 1 2 3 1 0; do echo "RVB_ROCK_CTAS=$c"; RVB_ROCK_CTAS=$c python bench.py --steps 100 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('  value %.3fM  ms %.4f  raycast %.4f  e2e %.3fM'%(d['value']/1e6,d['ms_per_step'],d['raycast_ms'],d['e2e']['value']/1e6))"; done
