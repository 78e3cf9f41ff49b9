cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2
(time timeout 1100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 --config 4 --genomes 10000 --no-cpu-baseline --max-seconds 1000) > gpurun_out/n8_config5.log 2>&1
tail -c 1500 gpurun_out/n8_config5.log
