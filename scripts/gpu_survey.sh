python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "span_many_tiles" 2>&1 | tail -15
