# usage: bash tools/gpu_sweep.sh "<chunk sizes>"   (GPU box; writes gpurun_out/bench_p_<chunk>.json)
mkdir -p gpurun_out
# (run the parity suite separately)
for cf in ${1:-4864}; do
FB200_CHUNK_FRAMES=$cf python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_p_$cf.json 2> gpurun_out/bench_p_$cf.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_p_$cf.json').read().strip().splitlines()[-1]); print($cf, 'value', round(d['value']/1e9,2), 'e2e', round(d['e2e']['value']/1e9,2), round(d['e2e']['ms_per_step'],2), round(d['e2e']['h2d_ms_per_step'],2), round(d['e2e']['d2h_ms_per_step'],2), {k: round(v,3) for k,v in d['kernel_ms_per_step'].items()}, round(d['ms_per_step'],3))"
tail -2 gpurun_out/bench_p_$cf.err
done
