#!/bin/bash
# final evidence run of round 2 (second session): GPU suite, smoke, full bench, launch list, ncu of the Polynomial path
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_suite.log 2>&1
timeout 120 python __graft_entry__.py --smoke > gpurun_out/r2b_smoke.log 2>&1
timeout 500 python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 900 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --no-configs --steps 2 --warmup 1 > gpurun_out/r2b_launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_poly1s|k_poly|k_plan|k_sfold" -s 4 -c 4 -f -o gpurun_out/r2b python tools/prof_front.py 72 2 > gpurun_out/r2b_ncu.log 2>&1
