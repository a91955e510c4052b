#!/usr/bin/env bash
# One GPU-box session of a round: parity tests, bench (own arm + reference arm), ncu launch list and one full capture of K1.
# usage (from the repo root, on the GPU box): bash tools/gpu_round.sh <tag>      -> gpurun_out/<tag>_*
tag=${1:-rX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 400 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_ncu_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --profiler-range > gpurun_out/${tag}_ncu_bench_stdout.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aas_attn -s 2 -c 1 -f -o gpurun_out/${tag}_attn \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --triplets 512 > gpurun_out/${tag}_ncu_full_stdout.log 2>&1
python tools/ncu_digest.py gpurun_out/${tag}_attn.ncu-rep --top 30 > gpurun_out/${tag}_attn_ncu_summary.txt 2>&1
head -30 gpurun_out/${tag}_attn_ncu_summary.txt
cut -c1-500 gpurun_out/${tag}_bench.json
