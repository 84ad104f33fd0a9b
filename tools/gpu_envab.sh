# A/B of environment switches with one build variant: tools/gpu_envab.sh "<nvcc extra>" "ENV1=a ENV2=b" "ENV1=c" ...
flags=$1; shift
FB200_NVCC_EXTRA="$flags" python flacenc_rs_b200/build.py --force > /dev/null 2> gpurun_out/build_envab.err || { echo "build failed"; tail -5 gpurun_out/build_envab.err; exit 1; }
for e in "$@"; do
  env $e python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[$flags | $e]', round(d['value']/1e9,2), 'e2e', round(d['e2e']['value']/1e9,2), {k: round(x,3) for k,x in d['kernel_ms_per_step'].items()}, round(d['ms_per_step'],3))"
done
