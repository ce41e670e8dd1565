#!/bin/bash
# Host-side batch build under different worker-pool settings (run on the GPU box).
for v in "PF_POOL_SPIN=4000" "PF_POOL_SPIN=0" "PF_POOL_SPIN=400000" "PF_POOL_SPAWN=1"; do
  echo "== $v"
  env $v PF_HOST_TIMING=1 python tools/e2e_breakdown.py 2>&1 | grep -E "scene pass|scene checks|meta pass|meta tables|host-only|PFSceneBuild" | tail -10
done
