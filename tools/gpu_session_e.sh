#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > $O/r2_gputest_e.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r2_gputest_e.log
bash tools/pv16_ab.sh
