#!/bin/bash
# compute-sanitizer over one small case per kernel family (tools/sanitize_cases.py); logs -> gpurun_out/r2_sanitizer_*.log
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 300 python tools/sanitize_cases.py > $O/r2_sanitizer_plain.log 2>&1; echo "plain rc=$?"; tail -3 $O/r2_sanitizer_plain.log
for tool in memcheck synccheck racecheck; do
  timeout ${SAN_TIMEOUT:-1500} compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize_cases.py > $O/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -c "^ok" $O/r2_sanitizer_$tool.log; grep "ERROR SUMMARY\|FAILURES\|RACECHECK SUMMARY" $O/r2_sanitizer_$tool.log | tail -3
done
