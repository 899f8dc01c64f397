python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --series 96 --no-cpu 2>&1 | python tools/benchline.py
ATSC_ENGINES=1 ATSC_WAVE_MI=48 python bench.py --steps 10 --warmup 3 --series 96 --no-cpu 2>&1 | python tools/benchline.py
python tools/class_profile.py --classes periodic,gauge --len 512 --series 6000 2>&1 | grep "=="
