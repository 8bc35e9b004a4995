#!/bin/bash
# compute-sanitizer over one small assembly of every kernel family (tools/sanitize_run.py): memcheck, racecheck (shared-memory
# hazards of the panel / staging buffers), synccheck.  Logs -> gpurun_out/r02_sanitizer_*.log (copied to profiles/).
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|all families|Error|hazard" gpurun_out/r02_sanitizer_$tool.log | sort | uniq -c | head -12
done
