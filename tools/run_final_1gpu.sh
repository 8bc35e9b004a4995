#!/bin/bash
# The single-GPU pass behind profiles/r01_bench_final3_c2_1gpu.json and profiles/r01_launches_final3_hexp2poisson_n128.csv:
#   gpurun --timeout 1200 -- 'bash tools/run_final_1gpu.sh'
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py > gpurun_out/r01_final3_c2.json 2> gpurun_out/r01_final3_c2.err; tail -c 200 gpurun_out/r01_final3_c2.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/r01_final3_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
