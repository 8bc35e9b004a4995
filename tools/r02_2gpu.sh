#!/bin/bash
# Round 2, two GPUs: the multi-GPU paths behind the C ABI and the strategy, then the 2-rank bench (weak scaling, all configs).
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/r02_2gpu.sh'
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_2gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_distributed.py tests/test_gpu_zzzzz_round2.py -q -p no:cacheprovider -x 2>&1 | tail -30 | tee gpurun_out/r02_tests_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
tail -3 gpurun_out/r02_bench_2gpu.err
head -c 1500 gpurun_out/r02_bench_2gpu.json
