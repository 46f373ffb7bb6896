timeout 900 python -m pytest tests/test_dist_gpu.py -q -x 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err || tail -20 gpurun_out/bench_n2.err
head -c 400 gpurun_out/bench_n2.json; echo
