cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/g_pytest.log 2>&1
tail -2 gpurun_out/g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
(time timeout 900 python bench.py) > gpurun_out/g_bench.log 2>&1
tail -c 200 gpurun_out/g_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'seed_scan_kernel|xdrop_warp_kernel' -c 4 -o gpurun_out/g_seed_full python tools/prof_trace.py 16 1,2 1 > gpurun_out/g_seed_ncu.log 2>&1
