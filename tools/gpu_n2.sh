cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q) > gpurun_out/n2_pytest.log 2>&1
tail -2 gpurun_out/n2_pytest.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline) > gpurun_out/n2_bench.log 2>&1
tail -c 300 gpurun_out/n2_bench.log
