#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r2_y_gputests.log 2>&1; echo "gputests rc=$?"; tail -5 $O/r2_y_gputests.log
