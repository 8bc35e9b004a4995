#!/bin/bash
# two GPUs: sharded CG + the strategy on several GPUs (drop-in), remaining round-2 tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_zzzzz_round2.py -q -p no:cacheprovider -k "sharded_cg or several_gpus or multi_device" 2>&1 | tail -30 | tee gpurun_out/r02_tests_2gpu_b.log
