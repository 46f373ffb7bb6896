timeout 900 python bench.py > gpurun_out/bench_default_final3.json 2> gpurun_out/bench_default_final3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_default_final3.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['library_baseline']['ms_per_step'])"; tail -2 gpurun_out/bench_default_final3.err
