#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 200 python tools/trace_pipeline.py 1536x1204x70 A > $O/r2d_trace_A.log 2>&1; echo "trace A rc=$?"; tail -18 $O/r2d_trace_A.log
timeout 200 python tools/trace_pipeline.py 1536x1204x70 B > $O/r2d_trace_B.log 2>&1; echo "trace B rc=$?"; tail -16 $O/r2d_trace_B.log
timeout 200 python tools/trace_pipeline.py 1536x1204x70 A "WT_REGS=80" > $O/r2d_trace_A_regs80.log 2>&1; echo "trace A regs rc=$?"; tail -18 $O/r2d_trace_A_regs80.log
