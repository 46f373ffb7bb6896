timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gather_rows or headline_size" 2>&1 | tail -3
for m in split fused split fused; do
B200TTA_ATTN_BWD=$m timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; print('$m', round(d['ms_per_step'],1), d['clocks']['sm_mhz'], {n:round(v['ms']) for n,v in k.items() if 'attn_bwd' in n})"
done
