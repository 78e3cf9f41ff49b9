cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 28 32767 24 34; do
PB_STRONG_SEED=$v timeout 300 python tools/prof_trace.py 16 2 3 2>&1 | grep -o "ms_seed': [0-9.]*" | sed "s/^/strong=$v /"
done
for v in 28 32767; do
(PB_STRONG_SEED=$v timeout 900 python bench.py --config 4 --no-cpu-baseline --steps 2 --workers 3) > gpurun_out/b_bench_s$v.log 2>&1; grep -o '"seconds": [0-9.]*' gpurun_out/b_bench_s$v.log | sed "s/^/strong=$v /"
done
