# e2e of the pipelined host path for chunk sizes x buffer sets: bash tools/gpu_nsets.sh "<chunks>" "<nsets>"
for cf in ${1:-2432}; do for ns in ${2:-3 4}; do
FB200_CHUNK_FRAMES=$cf FB200_NSETS=$ns python bench.py --steps 6 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunk $cf nsets $ns e2e', round(d['e2e']['value']/1e9,2), sorted(d['e2e']['ms_steps_rank0']), 'h2d', round(d['e2e']['h2d_ms_per_step'],2))"
done; done
