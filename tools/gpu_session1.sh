mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/s1_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
tail -5 gpurun_out/s1_pytest.log
timeout 400 python tools/perf_attn.py diffsim_b200/_lib/libds_base.so diffsim_b200/_lib/libdiffsim_b200.so diffsim_b200/_lib/libds_pvu.so > gpurun_out/s1_perf_ab.txt 2>&1
cat gpurun_out/s1_perf_ab.txt
timeout 300 python tools/bench_retrieval.py --images 2032 --out gpurun_out/s1_retrieval_1gpu.json > gpurun_out/s1_retrieval.log 2>&1
tail -3 gpurun_out/s1_retrieval.log
timeout 600 python bench.py > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
cat gpurun_out/s1_bench.json | cut -c1-600
