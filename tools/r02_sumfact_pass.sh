#!/bin/bash
# one GPU: parity + timing of the sum-factorisation variants, headline bench line, ncu of the default kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zzz_sumfact.py tests/test_gpu_parity.py -q -p no:cacheprovider --tb=short -x -k "sumfact or hex" 2>&1 | tail -3
timeout 900 python tools/sumfact_check.py 96 13 20 2>&1 | grep -v Warning | tee gpurun_out/r02_sumfact_warp_variants_b.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline 2>gpurun_out/r02_c2.err | tee gpurun_out/r02_c2.json | cut -c 1-400
bash tools/ncu_capture.sh r02_ncu_sumfact_warp assemble_sumfact --grid 64 --no-extra
