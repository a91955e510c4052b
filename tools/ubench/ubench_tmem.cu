// Micro-benchmark of tcgen05.ld / tcgen05.st + wait cost per warp (B200): clocks per (load xN + wait::ld) iteration.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tmem ubench_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define LDX(N, REGS) asm volatile("tcgen05.ld.sync.aligned.32x32b.x" #N ".b32 " REGS ", [%" #N "];"

__device__ __forceinline__ void ld8(uint32_t a, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t a, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t a, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                 "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                 "=r"(r[30]), "=r"(r[31]) : "r"(a) : "memory");
}
__device__ __forceinline__ void st8(uint32_t a, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MODE 0: ld x8 + wait; 1: ld x16 + wait; 2: ld x32 + wait; 3: 2 x (ld x16) then one wait; 4: 4 x (ld x16) then one wait;
// 5: ld x16 + wait + st x8 (no wait::st); 6: ld x16 + wait + st x8 + wait::st
template <int MODE>
__global__ void k(uint32_t* out, long long* clk, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (uint32_t)((it * 16) & 255);
    if (MODE == 0) { uint32_t r[8]; ld8(base + col, r); wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) acc ^= r[j]; }
    if (MODE == 1 || MODE == 5 || MODE == 6) { uint32_t r[16]; ld16(base + col, r); wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc ^= r[j];
      if (MODE >= 5) { uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = acc + j;
        st8(base + 256 + col, w); if (MODE == 6) wait_st(); } }
    if (MODE == 2) { uint32_t r[32]; ld32(base + col, r); wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= r[j]; }
    if (MODE == 3) { uint32_t r[16], q[16]; ld16(base + col, r); ld16(base + col + 16, q); wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc ^= r[j] ^ q[j]; }
    if (MODE == 4) { uint32_t r[16], q[16], s[16], t[16]; ld16(base + col, r); ld16(base + col + 16, q); ld16(base + col + 32, s); ld16(base + col + 48, t); wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc ^= r[j] ^ q[j] ^ s[j] ^ t[j]; }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int MODE>
void run(const char* name, int warps) {
  uint32_t* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  int iters = 4000;
  k<MODE><<<148, warps * 32>>>(out, clk, 10);
  k<MODE><<<148, warps * 32>>>(out, clk, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  printf("%-44s warps=%2d : %.1f clk/iter  (%s)\n", name, warps, (double)c / iters, cudaGetErrorString(e));
  cudaFree(out); cudaFree(clk);
}

int main() {
  for (int w = 4; w <= 16; w *= 2) {
    run<0>("ld x8 + wait::ld", w);
    run<1>("ld x16 + wait::ld", w);
    run<2>("ld x32 + wait::ld", w);
    run<3>("2 x ld x16, one wait::ld", w);
    run<4>("4 x ld x16, one wait::ld", w);
    run<5>("ld x16 + wait::ld + st x8", w);
    run<6>("ld x16 + wait::ld + st x8 + wait::st", w);
  }
  return 0;
}
