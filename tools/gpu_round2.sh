#!/usr/bin/env bash
# Round-2 single-GPU session: parity tests, bench (own arm + reference arm), ncu launch list of the bench, one full capture of
# K1, DRAM / L2 metrics of the all-pairs matrix, pipe metrics of a small-head-dim shape.
# usage (repo root, on the GPU box): bash tools/gpu_round2.sh <tag>      -> gpurun_out/<tag>_*
tag=${1:-r2final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -n 4 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 400 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_ncu_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-torch-reference --no-retrieval --profiler-range > gpurun_out/${tag}_ncu_bench_stdout.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aas_attn -s 2 -c 1 -f -o gpurun_out/${tag}_attn \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-torch-reference --no-retrieval --no-secondary --triplets 512 > gpurun_out/${tag}_ncu_full_stdout.log 2>&1
python tools/ncu_digest.py gpurun_out/${tag}_attn.ncu-rep --top 30 > gpurun_out/${tag}_attn_ncu_summary.txt 2>&1
head -n 24 gpurun_out/${tag}_attn_ncu_summary.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -k regex:aas_attn --csv --log-file gpurun_out/${tag}_ncu_matrix508.csv \
  python tools/bench_retrieval.py --images 508 --reps 1 > gpurun_out/${tag}_ncu_matrix_stdout.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:aas_attn --csv --log-file gpurun_out/${tag}_ncu_shapes.csv \
  python tools/perf_shapes.py > gpurun_out/${tag}_ncu_shapes_stdout.log 2>&1
cut -c1-600 gpurun_out/${tag}_bench.json
