set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_sw_gpu.py tests/test_search_gpu.py -m gpu -x -q) > gpurun_out/b_pytest.log 2>&1
tail -3 gpurun_out/b_pytest.log
rm -f gpurun_out/b_sw.log
timeout 300 python tools/bench_sw.py 1000000 5 >> gpurun_out/b_sw.log 2>&1
PB_LIB_PATH=$PWD/gpurun_alt/libpb_vote.so timeout 300 python tools/bench_sw.py 1000000 5 >> gpurun_out/b_sw.log 2>&1
PB_LIB_PATH=$PWD/gpurun_alt/libpb_nopair.so timeout 300 python tools/bench_sw.py 1000000 5 >> gpurun_out/b_sw.log 2>&1
cat gpurun_out/b_sw.log
