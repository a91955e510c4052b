#!/usr/bin/env bash
# round 2, call A: time the experimental CTA-pair attention kernel against the default one
tag=r2a
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
echo "== default" > gpurun_out/${tag}_perf_attn.txt
timeout 300 python tools/perf_attn.py >> gpurun_out/${tag}_perf_attn.txt 2>&1
echo "== attn_pair=1" >> gpurun_out/${tag}_perf_attn.txt
DIFFSIM_B200_DEBUG=attn_pair=1 timeout 300 python tools/perf_attn.py >> gpurun_out/${tag}_perf_attn.txt 2>&1
cat gpurun_out/${tag}_perf_attn.txt
DIFFSIM_B200_DEBUG=attn_pair=1 timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/${tag}_bench_pair.json 2> gpurun_out/${tag}_bench_pair.err
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
cut -c1-300 gpurun_out/${tag}_bench_pair.json; cut -c1-300 gpurun_out/${tag}_bench_default.json
DIFFSIM_B200_DEBUG=attn_pair=1 timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_pair.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest_pair.log
tail -15 gpurun_out/${tag}_pytest_pair.log
