cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for w in 1 2 3; do (time timeout 900 python bench.py --config 4 --no-cpu-baseline --workers $w --steps 2) > gpurun_out/b_bench_w$w.log 2>&1; grep -o '"config4": {"workload.\{0,700\}' gpurun_out/b_bench_w$w.log | cut -c1-900; done
