timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 2 --warmup 3 --lat-h 96 --lat-w 160 --bsa-sparsity 0.9375 --no-cpu-baseline > gpurun_out/bench_720p_bsa_n8.json 2> gpurun_out/bench_720p_bsa_n8.err || tail -20 gpurun_out/bench_720p_bsa_n8.err
head -c 300 gpurun_out/bench_720p_bsa_n8.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 2 --warmup 2 --method full --no-cpu-baseline > gpurun_out/bench_full_n8.json 2> gpurun_out/bench_full_n8.err || tail -20 gpurun_out/bench_full_n8.err
head -c 300 gpurun_out/bench_full_n8.json; echo
