timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
( time timeout 900 python bench.py > gpurun_out/bench_default_final.json 2> gpurun_out/bench_default_final.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_default_final.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"], d["roofline"]["frac"], d["library_baseline"]["ms_per_step"] if d.get("library_baseline") else None, d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
for k,v in d["kernel_ms_per_step"].items(): print("  ",k,v["ms"],v["n"],v.get("tflops"))
PY
timeout 900 python bench.py --method full --steps 2 --warmup 2 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_full2.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_full2.json')); print('full', d['ms_per_step'], {k:v['ms'] for k,v in d['kernel_ms_per_step'].items() if k in ('colsum','mt_sgd','mt_sumsq','ln_mod_bwd')})"
