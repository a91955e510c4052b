#!/usr/bin/env bash
# Multi-GPU session: NCCL tests, sharded all-pairs retrieval (bitwise check against 1 GPU), bench.py at N GPUs.
# usage (repo root, on an N-GPU box): bash tools/gpu_round_multi.sh <tag> <N>
tag=${1:-rX}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/${tag}_pytest_dist.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest_dist.log
tail -4 gpurun_out/${tag}_pytest_dist.log
timeout 600 python tools/bench_retrieval.py --images 2032 --out gpurun_out/${tag}_retrieval_1gpu.json > gpurun_out/${tag}_retrieval.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tools/bench_retrieval.py --images 2032 --check --out gpurun_out/${tag}_retrieval_${N}gpu.json >> gpurun_out/${tag}_retrieval.log 2>&1
grep -h workload gpurun_out/${tag}_retrieval.log | cut -c1-900
timeout 600 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 40 --warmup 3 > gpurun_out/${tag}_bench_${N}gpu.json 2>> gpurun_out/${tag}_bench.err
for f in gpurun_out/${tag}_bench_1gpu.json gpurun_out/${tag}_bench_${N}gpu.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['n_gpus'], round(d['value']), d['unit'], 'e2e', round(d['e2e']['value']) if d.get('e2e') else None, d['clocks']['sm_mhz'])"; done
tail -c 600 gpurun_out/${tag}_bench.err
