// What can the tensor pipe sustain on K1's OWN instruction mix?  A persistent kernel (one CTA per SM) whose single issuing
// thread replays the MMAs of K1's items at D = 160 -- per half: 10 x (M128 N128 K16, A and B from shared memory) for
// S = Q K^T, then 8 x [(M128 N160 K16, A from TMEM) + (M128 N16 K16, A from TMEM)] for O += P V and l += P 1 -- with static
// operands (no TMA traffic, no softmax, no epilogue; operands hold random fp16 so that the datapaths toggle).  One commit per
// item, at most two items in flight.  Modes: 0 K1 mix, 1 QK only, 2 PV only (no row-sum MMAs), 3 PV + row sums,
// 4 N = 256 SS MMAs (the GEMM-like shape, for scale), 5 K1 mix without the row-sum MMAs.
// Prints SM clocks per item, ms, useful TFLOP/s (K1's algorithmic 4 * 128 * 256 * 160 flops per item) and the average SM clock.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I diffsim_b200/csrc -o tools/ubench/ubench_k1_mma tools/ubench/ubench_k1_mma.cu
#include <cstdio>
#include <cstdlib>

#include "ds_ptx.cuh"

using namespace ds;

constexpr int kD = 160, kSubBytes = 64, kNSub = 5;          // D = 160: five 64-byte-swizzled sub-tiles of 32 elements
constexpr int kQBytes = kNSub * 128 * kSubBytes;            // 40 KB
constexpr int kTileBytes = kNSub * 128 * kSubBytes;         // one K or V half: 40 KB
constexpr int kSmem = 1024 + kQBytes + 4 * kTileBytes + 2048 + 256;

__device__ __forceinline__ void spin(int clk) {
  if (clk <= 0) return;
  const long long t = clock64();
  while (clock64() - t < clk) {}
}

// gap: SM clocks the issuing thread idles before each of the four bursts of an item (mode 0 only) -- how much per-burst
// overhead of the issuing thread does the tensor queue hide?
__global__ void __launch_bounds__(128, 1) k1_mma_kernel(int mode, int items, long long* clk_out, float* sink, int gap) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + kQBytes;                // K_A, V_A, K_B, V_B
  uint8_t* sOnes = sKV + 4 * kTileBytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sOnes + 2048);   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // pseudo-random fp16 in (-1, 1)
  uint32_t s = 1234567u + threadIdx.x * 7919u + blockIdx.x * 104729u;
  for (int i = threadIdx.x; i < (kQBytes + 4 * kTileBytes + 2048) / 4; i += blockDim.x) {
    s = s * 1664525u + 1013904223u;
    const float a = ((s >> 8) & 0xffff) / 32768.0f - 1.0f, b = ((s >> 16) & 0xffff) / 32768.0f - 1.0f;
    __half2 h = __floats2half2_rn(a, b);
    reinterpret_cast<uint32_t*>(smem)[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  // P operand region (TMEM columns [0, 128)): random packed fp16
  {
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    for (int c = 0; c < 128; c += 16) {
      uint32_t v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        s = s * 1664525u + 1013904223u;
        __half2 h = __floats2half2_rn(((s >> 8) & 0xff) / 256.0f, ((s >> 16) & 0xff) / 256.0f);
        v[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      tmem_st_x16(tmem + lane_addr + c, v);
    }
    tmem_wait_st();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();

  if (warp == 0 && elect_one()) {
    constexpr uint32_t SBO = 8 * kSubBytes;
    const uint64_t q_desc0 = umma_smem_desc(smem_u32(sQ), 16, SBO, UMMA_SW64);
    const uint64_t kv_desc0 = umma_smem_desc(smem_u32(sKV), 16, SBO, UMMA_SW64);
    const uint64_t v_desc0 = umma_smem_desc(smem_u32(sKV), 128 * kSubBytes, SBO, UMMA_SW64);
    const uint64_t ones_desc = umma_smem_desc(smem_u32(sOnes), 16, 1024, UMMA_SW128);
    const uint32_t idesc_qk = umma_idesc_f16(0, 128, 128, 0, 0);
    const uint32_t idesc_pv = umma_idesc_f16(0, 128, kD, 0, 1);
    const uint32_t idesc_l = umma_idesc_f16(0, 128, 16, 0, 0);
    const uint32_t idesc_256 = umma_idesc_f16(0, 128, 256, 0, 0);
    auto qk = [&](int h) {
      const uint64_t k_desc = kv_desc0 + (uint64_t)((2 * h * kTileBytes) >> 4);
#pragma unroll
      for (int kc = 0; kc < kD / 16; ++kc) {
        const int sub = kc / 2, off = (kc % 2) * 32;
        umma_f16_ss(tmem + 128, q_desc0 + (uint64_t)((sub * 128 * kSubBytes + off) >> 4),
                    k_desc + (uint64_t)((sub * 128 * kSubBytes + off) >> 4), idesc_qk, kc > 0);
      }
    };
    // TMEM: P (A operand of PV, valid fp16 throughout) [0, 128), S accumulators of both halves aliased at [128, 256) (never
    // read), O [256, 416), l [416, 432)
    auto pv = [&](int h, bool with_l) {
      const uint64_t v_desc = v_desc0 + (uint64_t)(((2 * h + 1) * kTileBytes) >> 4);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t a_tmem = tmem + (ks >> 1) * 32 + (ks & 1) * 8;
        umma_f16_ts(tmem + 256, a_tmem, v_desc + (uint64_t)((ks * 16 * kSubBytes) >> 4), idesc_pv, 1u);
        if (with_l) umma_f16_ts(tmem + 416, a_tmem, ones_desc, idesc_l, 1u);
      }
    };
    const long long t0 = clock64();
    for (int it = 0; it < items; ++it) {
      if (it >= 2) mbar_wait(&bar[it & 1], ((it - 2) >> 1) & 1);
      tc_fence_after_sync();
      switch (mode) {
        case 0: spin(gap); pv(0, true); spin(gap); qk(0); spin(gap); pv(1, true); spin(gap); qk(1); break;
        case 1: qk(0); qk(1); break;
        case 2: pv(0, false); pv(1, false); break;
        case 3: pv(0, true); pv(1, true); break;
        case 4:
#pragma unroll
          for (int kc = 0; kc < 20; ++kc) {
            const int sub = (kc % 10) / 2, off = (kc % 2) * 32;
            umma_f16_ss(tmem + 256, q_desc0 + (uint64_t)((sub * 128 * kSubBytes + off) >> 4),
                        kv_desc0 + (uint64_t)((sub * 128 * kSubBytes + off) >> 4), idesc_256, kc > 0);
          }
          break;
        default: pv(0, false); qk(0); pv(1, false); qk(1); break;
      }
      umma_commit(&bar[it & 1]);
    }
    for (int it = items - 2 > 0 ? items - 2 : 0; it < items; ++it) mbar_wait(&bar[it & 1], (it >> 1) & 1);
    const long long t1 = clock64();
    if (blockIdx.x == 0) *clk_out = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
  if (sink && threadIdx.x == 0 && blockIdx.x == 0) sink[0] = 1.f;
}

int main(int argc, char** argv) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* clk;
  float* sink;
  cudaMalloc(&clk, 8);
  cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(k1_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  const char* names[6] = {"K1 mix (QK + PV + row sums)", "QK only (N=128 SS)", "PV only (N=160 TS)", "PV + row sums",
                          "N=256 SS (GEMM-like)", "K1 mix without row sums"};
  // flops actually issued per item and the K1-useful flops per item
  const double issued[6] = {2.0 * (2 * 128. * 128 * 160 + 2 * 128. * 176 * 128), 2.0 * 2 * 128. * 128 * 160,
                            2.0 * 2 * 128. * 160 * 128, 2.0 * 2 * 128. * 176 * 128, 2.0 * 20 * 128. * 256 * 16,
                            2.0 * (2 * 128. * 128 * 160 + 2 * 128. * 160 * 128)};
  const double useful[6] = {4.0 * 128 * 256 * 160, 2.0 * 128 * 256 * 160, 2.0 * 128 * 256 * 160, 2.0 * 128 * 256 * 160,
                            issued[4], 4.0 * 128 * 256 * 160};
  const int reps = argc > 1 ? atoi(argv[1]) : 3;
  for (int rep = 0; rep < reps; ++rep)
    for (int mode = 0; mode < 6; ++mode) {
      const int items = 200000;   // ~0.3 - 0.6 s per launch: long enough for the power governor to settle
      k1_mma_kernel<<<sms, 128, kSmem>>>(mode, 2000, clk, sink, 0);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0);
      k1_mma_kernel<<<sms, 128, kSmem>>>(mode, items, clk, sink, 0);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) {
        printf("mode %d: %s\n", mode, cudaGetErrorString(err));
        return 1;
      }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      long long c = 0;
      cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
      printf("[rep %d] %-30s %7.0f clk/item  %8.2f ms  issued %6.0f TFLOP/s  K1-useful %6.0f TFLOP/s  avg SM clock %.2f GHz\n", rep,
             names[mode], (double)c / items, ms, issued[mode] * items * sms / (ms * 1e-3) / 1e12,
             useful[mode] * items * sms / (ms * 1e-3) / 1e12, (double)c / (ms * 1e6));
      fflush(stdout);
    }
  // queue depth probe: the K1 mix with the issuing thread idling `gap` clocks before every burst
  for (int gap : {0, 50, 100, 200, 400, 800}) {
    const int items = 50000;
    k1_mma_kernel<<<sms, 128, kSmem>>>(0, items, clk, sink, gap);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    printf("[gap %4d clk before each burst] %7.0f clk/item (+%.0f over 4 x gap = %d)\n", gap, (double)c / items,
           (double)c / items - 2881.0, 4 * gap);
  }
  return 0;
}
