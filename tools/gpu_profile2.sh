# ncu captures for profiles/: full set (with source counters) of the heaviest kernels + launch list of one bench step
tag=${1:-r2}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"fb_ka_plan|fb_kp_pack|fb_k1_analyze|fb_k0_ingest" -s 0 -c 4 -o gpurun_out/prof_${tag} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_${tag}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
tail -2 gpurun_out/ncu_${tag}.log | cut -c1-200
