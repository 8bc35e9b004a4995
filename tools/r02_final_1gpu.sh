#!/bin/bash
# one GPU, end of round 2: ncu launch list (time + DRAM bytes of every launch of two bench steps), ncu --set full of the elasticity
# kernel, compute-sanitizer over every kernel family, the driver-format bench line, the whole GPU suite
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c2_n128.csv \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/r02_launches_c2_n128.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c5_n81.csv \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline --no-e2e --phys elasticity --grid 81 > gpurun_out/r02_launches_c5_n81.log 2>&1
bash tools/ncu_capture.sh r02_ncu_elast_pair assemble_gram_warp --grid 48 --phys elasticity --no-extra
bash tools/r02_sanitizer.sh
( time python bench.py ) > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -4 gpurun_out/r02_bench_final.err; cut -c 1-200 gpurun_out/r02_bench_final.json
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short ) > gpurun_out/r02_gpu_suite_final.log 2>&1
tail -6 gpurun_out/r02_gpu_suite_final.log
