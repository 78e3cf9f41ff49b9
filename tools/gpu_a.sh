set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/a_pytest.log 2>&1
(time timeout 600 python bench.py) > gpurun_out/a_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 2 --warmup 1 --genomes 32 --c3-genomes 2 --no-cpu-baseline > gpurun_out/a_bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sw_band_trace_kernel|sw_walk_kernel|band_bounds' -c 8 -o gpurun_out/a_trace_full python tools/prof_trace.py 16 1,2 1 > gpurun_out/a_trace_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'seed_scan_kernel|xdrop_warp_kernel' -c 4 -o gpurun_out/a_seed_full python tools/prof_trace.py 16 1,2 1 > gpurun_out/a_seed_ncu.log 2>&1
ls -la gpurun_out
