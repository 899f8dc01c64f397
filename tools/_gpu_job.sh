python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python tools/benchline.py
