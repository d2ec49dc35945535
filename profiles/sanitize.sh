#!/bin/bash
# profiles/sanitize.sh -- compute-sanitizer over __graft_entry__.smoke() (trna: sieve kernel -> enumeration kernel;
# score.1: fused kernel; both checked against the oracle inside smoke()).  Logs go to gpurun_out/, summaries to profiles/.
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" gpurun_out/sanitize_$tool.log | tail -4
done
