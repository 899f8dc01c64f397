python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/class_profile.py --series 48 --classes noisy,periodic 2>&1 | grep "=="
python tools/class_profile.py --series 48 --classes periodic,util --comp fft 2>&1 | grep "==\|frame0"
python tools/class_profile.py --series 1 --classes periodic --comp fft --error 1 2>&1 | grep "==\|frame0"
python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | python tools/benchline.py
