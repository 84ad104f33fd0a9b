# A/B of one environment switch: tools/gpu_ab.sh VAR val1 val2 ...   (kernel breakdown of the default bench)
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
  tail -1 gpurun_out/bench_ab_$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$var=$v', round(d['value']/1e9,2), {k: round(x,3) for k,x in d['kernel_ms_per_step'].items()})"
done
