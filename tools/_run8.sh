cd $GRAFT_REPO_ROOT
timeout 160 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
grep -o '"value": [0-9.]*' gpurun_out/bench_n8.json | head -2; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_n8.json | head -1
tail -3 gpurun_out/bench_n8.err
