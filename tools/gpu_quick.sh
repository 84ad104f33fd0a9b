# quick kernel-time check at a chunk-sized batch (600 s) and at the full hour
for sec in 600 3600; do
python bench.py --seconds $sec --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_q_$sec.json 2> gpurun_out/bench_q_$sec.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_q_$sec.json').read()); print($sec, 'value', round(d['value']/1e9,2), 'e2e', round(d['e2e']['value']/1e9,2), round(d['e2e']['ms_per_step'],2), {k: round(v,3) for k,v in d['kernel_ms_per_step'].items()}, round(d['ms_per_step'],3))"
tail -2 gpurun_out/bench_q_$sec.err
done
