# 8-GPU evidence run (gpurun --gpus 8 -- bash tools/run_n8.sh): torchrun bench, multi-device context tests, single-process bench
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 < /dev/null > gpurun_out/fin_n8.json 2> gpurun_out/fin_n8.err
timeout 300 python -m pytest tests/test_gpu_multidev.py -q -m gpu < /dev/null > gpurun_out/fin_multidev8.log 2>&1
timeout 400 python bench.py --gpus 8 --single-process --steps 5 --warmup 3 < /dev/null > gpurun_out/fin_sp8.json 2> gpurun_out/fin_sp8.err
