python -m pytest tests -m gpu -x -q 2>&1 | tail -3
( time python bench.py > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err ) 2>&1 | grep real
cat gpurun_out/bench_r1_n1.json | python tools/benchline.py; tail -3 gpurun_out/bench_r1_n1.err
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err ) 2>&1 | grep real
cat gpurun_out/bench_r1_ref.json | cut -c1-400
