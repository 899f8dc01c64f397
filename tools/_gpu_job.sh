( time python -m pytest tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -15 ) 2>&1 | tail -20
