# e2e A/B of one environment switch: bash tools/gpu_env.sh VAR "v1 v2 ..." [repeats]
var=$1
for rep in $(seq ${3:-1}); do for v in $2; do
env $var=$v python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$var=$v value', round(d['value']/1e9,2), 'e2e', round(d['e2e']['value']/1e9,2), sorted(d['e2e']['ms_steps_rank0']), 'h2d', round(d['e2e']['h2d_ms_per_step'],2))"
done; done
