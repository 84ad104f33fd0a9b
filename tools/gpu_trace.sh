# device timeline of the pipelined host path (last chunks) for a few chunk sizes
for cf in ${1:-4864}; do
echo "== chunk $cf"
FB200_TRACE=1 FB200_CHUNK_FRAMES=$cf python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "fb200 trace" | tail -24
done
