#!/bin/bash
# one GPU, last check of the committed state: smoke(), the drop-in tests, the driver-format bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warning | tee gpurun_out/r02_smoke.log
timeout 600 python -m pytest tests/test_gpu_dropin.py -q -p no:cacheprovider --tb=short -x 2>&1 | tail -3
( time python bench.py ) > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -4 gpurun_out/r02_bench_final.err; cut -c 1-160 gpurun_out/r02_bench_final.json
