timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -x 2>&1 | tail -15
for ov in 1 0; do
B200TTA_OVERLAP_ALLREDUCE=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$ov bench.py --gpus 2 --steps 3 --warmup 2 --method full > gpurun_out/bench_full_n2_ov$ov.json 2> gpurun_out/bench_full_n2_ov$ov.err; python -c "
import json; d=json.load(open('gpurun_out/bench_full_n2_ov$ov.json')); print('overlap=$ov', d['ms_per_step'], d['value'], d['loss_first_last'] if 'loss_first_last' in d else '')"; tail -2 gpurun_out/bench_full_n2_ov$ov.err
done
