python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/b14.json 2>gpurun_out/b14.err; cat gpurun_out/b14.json | python tools/benchline.py; tail -2 gpurun_out/b14.err; python -c "
import json; d=json.load(open('gpurun_out/b14.json')); print(d['decompress'])"
