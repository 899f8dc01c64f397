python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/class_profile.py --series 48 --classes saw,noisy,steps 2>&1 | grep "=="
python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python tools/benchline.py
