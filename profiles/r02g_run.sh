#!/bin/bash
# round 2: full GPU suite (host layer incl. RM3 constrained, sweep, kinematics; multi-device handle), smoke
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02g_pytest.log
tail -25 gpurun_out/r02g_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02g_smoke.log 2>&1; tail -3 gpurun_out/r02g_smoke.log
