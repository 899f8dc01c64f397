python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -30
