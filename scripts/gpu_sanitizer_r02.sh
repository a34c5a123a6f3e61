#!/bin/bash
# GPU box: compute-sanitizer over smoke() (one small train step through every kernel of the path) with the three tools SURVEY.md
# section 5 asks for.  Summaries (error counts + first reports) land in gpurun_out/sanitizer_r02_<tool>.log.
set -u
mkdir -p gpurun_out
for tool in memcheck initcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_r02_$tool.full 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_r02_$tool.full
  { echo "# compute-sanitizer --tool $tool  python -c 'import __graft_entry__ as g; g.smoke()'   ($(date -u +%FT%TZ))";
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke:|^exit|Error|error" gpurun_out/sanitizer_r02_$tool.full | head -40;
    echo "--- first lines of reports";
    grep -E "=========" gpurun_out/sanitizer_r02_$tool.full | head -60; } > gpurun_out/sanitizer_r02_$tool.log
  tail -3 gpurun_out/sanitizer_r02_$tool.log
done
