timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 2 --warmup 3 --lat-h 96 --lat-w 160 --bsa-sparsity 0.9375 --no-cpu-baseline > gpurun_out/bench_720p_bsa_n8_v2.json 2> gpurun_out/bench_720p_bsa_n8_v2.err || tail -20 gpurun_out/bench_720p_bsa_n8_v2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_720p_bsa_n8_v2.json')); print(d['value'], d['ms_per_step'], d['clocks']); print(sum(v['ms'] for v in d['kernel_ms_per_step'].values()))"
