for i in 1 2 3; do timeout 300 python -m pytest tests/test_text_encoder_gpu.py -m gpu -q -k "without_mask" --tb=short 2>&1 | grep -E "assert|Error|passed|failed" | head -5; done
