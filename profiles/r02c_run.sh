#!/bin/bash
mkdir -p gpurun_out
for ns in 2 1; do for la in 2 1; do echo "HC_RB_STREAMS=$ns rad_lookahead=$la"; HC_RB_STREAMS=$ns timeout 300 python profiles/dbg_wrong_hint.py $la; done; done > gpurun_out/r02c_dbg.log 2>&1
cat gpurun_out/r02c_dbg.log
