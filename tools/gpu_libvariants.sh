# A/B of prebuilt library variants (tools/variants/v*.so + v*.txt with their flags): kernel breakdown of the default bench
cp flacenc_rs_b200/csrc/libflacenc_b200.so /tmp/lib_default.so
for so in tools/variants/v*.so; do
  cp $so flacenc_rs_b200/csrc/libflacenc_b200.so
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[$(cat ${so%.so}.txt)]', round(d['value']/1e9,2), {k: round(x,3) for k,x in d['kernel_ms_per_step'].items()}, round(d['ms_per_step'],3))"
done
cp /tmp/lib_default.so flacenc_rs_b200/csrc/libflacenc_b200.so
