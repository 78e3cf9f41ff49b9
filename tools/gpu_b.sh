set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_search_gpu.py tests/test_real_genomes.py tests/test_real_slice.py -m gpu -x -q) > gpurun_out/b_pytest.log 2>&1
tail -3 gpurun_out/b_pytest.log
timeout 300 python tools/prof_trace.py 16 1,2 2 > gpurun_out/b_search.log 2>&1
cat gpurun_out/b_search.log
