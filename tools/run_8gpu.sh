#!/bin/bash
# eight GPUs: the multi-GPU tests (2 and 4 devices behind the C ABI / the strategy, NCCL ranks), then the 8-rank bench (weak scaling, all configs)
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/run_8gpu.sh'
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_8gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_distributed.py tests/test_gpu_zzzzz_round2.py -q -p no:cacheprovider -k "sharded or several_gpus or multi" 2>&1 | tail -8 | tee gpurun_out/r02_tests_8gpu.log
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
tail -2 gpurun_out/r02_bench_8gpu.err | cut -c 1-300
head -c 600 gpurun_out/r02_bench_8gpu.json
