timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err; tail -c 600 gpurun_out/bench_n2_final.json; tail -3 gpurun_out/bench_n2_final.err
