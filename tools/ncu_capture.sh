#!/bin/bash
# usage: tools/ncu_capture.sh <tag> <kernel regex> <bench args...>
# One `ncu --set full` capture of the dominant kernel (second launch), exported on the GPU box to small CSV files
# (gpurun copies back at most 64 MiB): gpurun_out/<tag>_raw.csv (all metrics) and gpurun_out/<tag>_sass.csv (per-instruction
# stall samples).  The .ncu-rep itself is deleted.
tag=$1; regex=$2; shift 2
ncu --set full --clock-control none --import-source on -k regex:$regex -s 1 -c 1 -o gpurun_out/$tag python bench.py "$@" --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/$tag.log 2>&1
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/$tag.ncu-rep --page source --csv 2>/dev/null | cut -d, -f1-6,30-50 > gpurun_out/${tag}_sass.csv
rm -f gpurun_out/$tag.ncu-rep
ls -la gpurun_out/${tag}_raw.csv gpurun_out/${tag}_sass.csv
