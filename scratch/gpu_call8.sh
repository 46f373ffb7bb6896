nproc; free -g | head -2
( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | tail -3; tail -3 gpurun_out/bench_default.err; head -c 600 gpurun_out/bench_default.json; echo
for m in delta_a delta_b delta_c norm_tune film; do
  timeout 600 python bench.py --method $m --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err || tail -5 gpurun_out/bench_$m.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/bench_$m.json')); print('$m', d['ms_per_step'], d['e2e']['value'], d['config']['adapter_params'], d['loss_first_last'])
except Exception as e: print('$m failed', e)"
done
( time timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2>&1 | tail -3; head -c 1500 gpurun_out/bench_reference.json; echo
