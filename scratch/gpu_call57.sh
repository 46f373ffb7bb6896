timeout 300 python -m pytest tests/test_cli_gpu.py -m gpu -q -k "full" --tb=short 2>&1 | tail -12
timeout 120 python lora_experiment/scripts/run_full_tta.py --output-dir /tmp/full_cli --synthetic --model tiny --latent-hw 32,32 --tta-total-frames 17 --tta-context-frames 5 --num-steps 2 --max-videos 1 --es-disable 2>&1 | tail -3; python -c "
import json; s=json.load(open('/tmp/full_cli/summary.json')); print(s['method'], s['num_successful'], s['results'][0]['losses'])"
