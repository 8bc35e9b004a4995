#!/bin/bash
# one GPU: the whole GPU suite, then the driver-format bench (both arms) with wall-clock times
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short -x ) > gpurun_out/r02_gpu_suite_d.log 2>&1
tail -6 gpurun_out/r02_gpu_suite_d.log
( time python bench.py --impl reference ) > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
tail -4 gpurun_out/r02_bench_ref.err; cut -c 1-300 gpurun_out/r02_bench_ref.json
( time python bench.py ) > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err
tail -4 gpurun_out/r02_bench_d.err; cut -c 1-200 gpurun_out/r02_bench_d.json
