cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/f_pytest.log 2>&1
tail -3 gpurun_out/f_pytest.log
(time timeout 900 python bench.py) > gpurun_out/f_bench.log 2>&1
tail -c 200 gpurun_out/f_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 1 --genomes 32 --c3-genomes 2 --no-cpu-baseline --workers 1 > gpurun_out/f_bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sw_kernel -c 2 -o gpurun_out/f_sw_full python tools/bench_sw.py 1000000 1 > gpurun_out/f_sw_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'seed_scan_kernel|xdrop_warp_kernel' -c 4 -o gpurun_out/f_seed_full python tools/prof_trace.py 16 1,2 1 > gpurun_out/f_seed_ncu.log 2>&1
(time timeout 900 python bench.py --config 3 --no-cpu-baseline --steps 2) > gpurun_out/f_bench_c3.log 2>&1
grep -o '"ladder_seconds": [0-9.]*\|"genes_clustered_per_s": [0-9.]*' gpurun_out/f_bench_c3.log
ls -la gpurun_out | tail -8
