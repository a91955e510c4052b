#!/usr/bin/env bash
# quick K1 iteration: parity (one file), burst timing, timeline.  usage: bash tools/gpu_quick.sh <tag> [libs...]
tag=${1:-q}; shift
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/${tag}_pytest_parity.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest_parity.log
tail -n 4 gpurun_out/${tag}_pytest_parity.log
timeout 600 python tools/perf_attn.py "$@" > gpurun_out/${tag}_perf_attn.txt 2>&1; cat gpurun_out/${tag}_perf_attn.txt
if [ -f diffsim_b200/_lib/libds_trace.so ]; then
DIFFSIM_B200_LIB=$PWD/diffsim_b200/_lib/libds_trace.so TR_LO=60000 TR_HI=75000 timeout 120 python tools/trace_attn.py > gpurun_out/${tag}_trace.txt 2>&1; tail -n 1 gpurun_out/${tag}_trace.txt
fi
