# ncu captures for profiles/: full set of the two heaviest kernels + launch list of one bench step (GPU box)
tag=${1:-r1}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"fb_ka_plan|fb_kp_pack|fb_k1_analyze" -s 0 -c 3 -o gpurun_out/prof_${tag} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_${tag}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
tail -2 gpurun_out/ncu_${tag}.log | cut -c1-200
