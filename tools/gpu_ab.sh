# A/B of an environment switch: bash tools/gpu_ab.sh VAR  (runs bench with VAR unset and VAR=1)
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in 0 1; do
if [ $v = 0 ]; then E="env -u $1"; else E="env $1=1"; fi
$E python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_ab_$v.json').read()); print('$1=$v', 'value', round(d['value']/1e9,2), 'e2e', round(d['e2e']['value']/1e9,2), {k: round(v,3) for k,v in d['kernel_ms_per_step'].items()}, round(d['ms_per_step'],3))"
tail -2 gpurun_out/bench_ab_$v.err
done
