cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_search_gpu.py tests/test_real_genomes.py tests/test_real_slice.py tests/test_clust_gpu.py tests/test_uberblast_gpu.py -m gpu -x -q) > gpurun_out/b_pytest.log 2>&1
tail -2 gpurun_out/b_pytest.log
timeout 300 python tools/prof_trace.py 16 1,2 3 2>&1 | grep -o "mode [12]\|ms_seed': [0-9.]*"
