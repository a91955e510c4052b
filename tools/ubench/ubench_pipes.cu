// Micro-benchmarks of the SM pipes the softmax warps live on (B200): MUFU.EX2, F2FP, FFMA2, FMNMX3 and a
// softmax-like mix, per warp count.  Prints SM clocks per warp-instruction per SMSP.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(float* out, long long* clk, int iters) {
  float x[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = (threadIdx.x * 0.001f + j * 0.01f) - 3.0f;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // 32 independent MUFU.EX2
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = ex2(x[j]);
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = x[j] * 0.5f - 3.0f;   // keep in range (32 FFMA)
    } else if (MODE == 1) {   // 16 F2FP packs + 32 ffma
#pragma unroll
      for (int j = 0; j < 32; j += 2) { __half2 h = __floats2half2_rn(x[j], x[j + 1]); acc ^= *reinterpret_cast<uint32_t*>(&h); }
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = x[j] * 0.5f - 3.0f;
    } else if (MODE == 2) {   // softmax-like chunk: 16 max3, 16 ffma2-ish (32 ffma), 32 ex2, 16 pack
      float m = -1e30f;
#pragma unroll
      for (int j = 0; j < 32; ++j) m = fmaxf(m, x[j]);
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = ex2(x[j] * 1.1f - m);
#pragma unroll
      for (int j = 0; j < 32; j += 2) { __half2 h = __floats2half2_rn(x[j], x[j + 1]); acc ^= *reinterpret_cast<uint32_t*>(&h); }
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = x[j] * 0.5f - 3.0f + (float)(acc & 1);
    } else if (MODE == 4) {   // 16 ex2.approx.ftz.f16x2 (= 32 exponentials) + 32 ffma: does the packed form double the MUFU rate?
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        __half2 h = __floats2half2_rn(x[j], x[j + 1]);
        uint32_t u = *reinterpret_cast<uint32_t*>(&h), r;
        asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(u));
        acc ^= r;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = x[j] * 0.5f - 3.0f + (float)(acc & 1);
    } else if (MODE == 3) {   // only 32 ffma (baseline for modes 0/1)
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = x[j] * 0.5f - 3.0f;
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_smsp) {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  int iters = 2000, threads = warps_per_smsp * 4 * 32;
  k<MODE><<<148, threads>>>(out, clk, 10);
  k<MODE><<<148, threads>>>(out, clk, iters);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  printf("%-28s warps/SMSP=%d : %.1f clk per iteration per warp-slot (%.1f clk/iter / warps)\n", name, warps_per_smsp, (double)c / iters, (double)c / iters / warps_per_smsp);
  cudaFree(out); cudaFree(clk);
}

int main() {
  for (int w = 1; w <= 4; w *= 2) {
    run<3>("32 FFMA", w);
    run<0>("32 MUFU.EX2 + 32 FFMA", w);
    run<1>("16 F2FP.PACK + 32 FFMA", w);
    run<2>("softmax-like chunk", w);
    run<4>("16 F2FP + 16 EX2.F16x2 + 32 FFMA", w);
  }
  return 0;
}
