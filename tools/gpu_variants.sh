# A/B of build variants: tools/gpu_variants.sh "<nvcc extra flags 1>" "<flags 2>" ...  (kernel breakdown of the default bench)
i=0
for flags in "$@"; do
  FB200_NVCC_EXTRA="$flags" python flacenc_rs_b200/build.py --force > /dev/null 2> gpurun_out/build_var_$i.err || { echo "build failed: $flags"; tail -5 gpurun_out/build_var_$i.err; continue; }
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_var_$i.json 2> gpurun_out/bench_var_$i.err
  tail -1 gpurun_out/bench_var_$i.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[$flags]', round(d['value']/1e9,2), 'e2e', round(d['e2e']['value']/1e9,2), {k: round(x,3) for k,x in d['kernel_ms_per_step'].items()}, round(d['ms_per_step'],3))"
  i=$((i+1))
done
