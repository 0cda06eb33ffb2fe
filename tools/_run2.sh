cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -x -q 2>&1 | tail -5
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
grep -o '"value": [0-9.]*' gpurun_out/bench_n2.json | head -2; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_n2.json | head -1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --parallel items > gpurun_out/bench_items_n2.json 2> gpurun_out/bench_items_n2.err
grep -o '"value": [0-9.]*' gpurun_out/bench_items_n2.json | head -2; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_items_n2.json | head -1
tail -3 gpurun_out/bench_n2.err
