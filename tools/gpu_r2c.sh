#!/usr/bin/env bash
# round 2: K1 restructure check -- parity first (short timeout: a hung pipeline traps via the mbarrier watchdog), then speed
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/${tag}_pytest_parity.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest_parity.log
tail -15 gpurun_out/${tag}_pytest_parity.log
timeout 300 python tools/perf_attn.py > gpurun_out/${tag}_perf_attn.txt 2>&1; cat gpurun_out/${tag}_perf_attn.txt
DIFFSIM_B200_LIB=$PWD/diffsim_b200/_lib/libds_trace.so TR_LO=60000 TR_HI=75000 timeout 120 python tools/trace_attn.py > gpurun_out/${tag}_trace.txt 2>&1; tail -n 2 gpurun_out/${tag}_trace.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest_all.log
tail -n 5 gpurun_out/${tag}_pytest_all.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-200 gpurun_out/${tag}_bench.json
