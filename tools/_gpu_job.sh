python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for r in 0 1; do python bench.py --steps 20 --warmup 3 --no-cpu --as-rank $r 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), {k:round(v,3) for k,v in r['kernel_ms_per_step'].items()}); print({k:round(v,3) for k,v in r['kernel_ms_per_step_one_engine'].items()}, d['device_ms_per_step']['compress'])"; done
python tools/class_profile.py --series 48 --classes steps,saw,gauge 2>&1 | grep "=="
