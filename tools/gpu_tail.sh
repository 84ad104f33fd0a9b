# e2e of the pipelined host path for chunk sizes x tail steps: bash tools/gpu_tail.sh "<chunks>" "<tail steps>"
for cf in ${1:-2432}; do for ts in ${2:-0 3}; do
FB200_CHUNK_FRAMES=$cf FB200_TAIL_STEPS=$ts python bench.py --steps 6 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunk $cf tail $ts e2e', round(d['e2e']['value']/1e9,2), sorted(d['e2e']['ms_steps_rank0']), 'h2d', round(d['e2e']['h2d_ms_per_step'],2))"
done; done
