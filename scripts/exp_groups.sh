for g in 1 2 3 4; do echo "HOST_GROUPS=$g"; B200DDSP_HOST_GROUPS=$g python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('resident ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['roofline']['stage_ms'])"; done
