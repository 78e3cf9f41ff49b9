cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/b_pytest.log 2>&1
tail -3 gpurun_out/b_pytest.log
for i in 1 2; do PB_DEBUG_TIMING=1 timeout 300 python tools/prof_trace.py 16 1,2 3 2>&1 | grep "device:" | cut -c1-200; done
timeout 300 python tools/bench_sw.py 1000000 5
