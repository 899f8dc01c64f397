python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | python tools/benchline.py
ATSC_ENGINES=1 ATSC_WAVE_MI=128 python bench.py --steps 10 --warmup 3 --series 96 --no-cpu 2>&1 | python tools/benchline.py
