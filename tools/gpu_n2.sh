cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --config 4 --no-cpu-baseline --workers 3) > gpurun_out/n2_bench_dbg.log 2>&1
grep -o '"seconds": [0-9.]*\|"rank0_seconds[a-z_]*": [0-9.]*\|"allgather_verified": [a-z]*' gpurun_out/n2_bench_dbg.log
