#!/usr/bin/env bash
# Round-2 multi-GPU session: NCCL tests, sharded all-pairs retrieval (bitwise check against 1 GPU), bench.py at N GPUs (own arm;
# the reference arm runs on rank 0 only).   usage (repo root, on an N-GPU box): bash tools/gpu_round_multi2.sh <tag> <N>
tag=${1:-rX}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/${tag}_pytest_dist.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest_dist.log
tail -n 3 gpurun_out/${tag}_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tools/bench_retrieval.py --images 2032 --check --out gpurun_out/${tag}_retrieval_${N}gpu.json > gpurun_out/${tag}_retrieval.log 2>&1
grep -h workload gpurun_out/${tag}_retrieval.log | cut -c1-700
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 40 --warmup 3 > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench_${N}gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value']), d['unit'], 'K1', round(d['roofline']['achieved']), 'e2e', round(d['e2e']['value']), d['e2e'].get('h2d_ceiling',{}).get('gbs_all_gpus'), d['e2e'].get('h2d_frac_of_ceiling'), 'retrieval', d.get('retrieval',{}).get('ms_total'), d['clocks']['sm_mhz'])"
tail -c 400 gpurun_out/${tag}_bench.err
