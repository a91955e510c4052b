# usage: bash tools/gpu_ab.sh lib1.so lib2.so ...   A/B of attention-kernel builds: burst (clocks per item, ms, TFLOP/s over
# 20 ms) and SUSTAINED (bench.py's 40 device-resident steps of 2048 triplets, power-capped regime)
mkdir -p gpurun_out
timeout 600 python tools/perf_attn.py "$@" > gpurun_out/perf_ab.txt 2>&1
cat gpurun_out/perf_ab.txt
for rep in 1 2; do for lib in "$@"; do
  echo -n "[sustained $(basename $lib) rep$rep] "
  DIFFSIM_B200_LIB=$(realpath $lib) timeout 300 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), 'pairs/s', round(d['roofline']['achieved'],1), 'TFLOP/s', d['clocks']['sm_mhz'], 'MHz', d['clocks']['power_w_max'], 'W')"
done; done | tee gpurun_out/sustained_ab.txt
