#!/usr/bin/env bash
# usage: bash tools/gpu_sustained.sh <tag> lib1.so lib2.so ...   sustained (power-capped) A/B: bench.py's 40 x 2048-triplet steps per library
tag=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do for lib in "$@"; do
  echo -n "[sustained $(basename $lib) rep$rep] "
  DIFFSIM_B200_LIB=$(realpath $lib) timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-secondary 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), 'pairs/s', round(d['roofline']['achieved'],1), 'TFLOP/s', d['clocks']['sm_mhz'], 'MHz', d['clocks']['power_w_max'], 'W')"
done; done | tee gpurun_out/${tag}_sustained.txt
