// Correctness probe for the two mechanisms the cta_group::2 form of the attention kernel
// rests on, in isolation.  Run on a B200 at the end of round 1: max |err| 4.1e-6 against |O| up to 14.7 -> all three
// questions below are answered with yes (profiles/r1z_ubench_pair_pv.txt).
//
//   O[256 x 160] (fp32) = P[256 x 128] (fp16, read from TENSOR MEMORY: the TS form of tcgen05.mma) . V[128 x 160] (fp16)
//
// as ONE CTA pair: each CTA holds its 128 rows of P in its own TMEM (packed 16-bit pairs, the layout the softmax warps of
// aas_attn_kernel write: the 16-row k step ks lives at columns (ks >> 1) * 32 + (ks & 1) * 8), and 80 of V's 160 columns
// in shared memory -- the N split of the B operand -- as five MN-major sub-tiles of 16 columns x 128 rows with the
// 32-BYTE swizzle (80 is not a multiple of the 32-column sub-tile the 1-CTA kernel uses at D = 160).  The leader issues
// eight M = 256, N = 160, K = 16 MMAs; both CTAs read their 128 x 160 accumulator rows back and the host compares with a
// double-precision product.
//
// Questions it answers: (1) does kind::f16 cta_group::2 accept A from TMEM with each CTA supplying its own half of M;
// (2) is the (LBO = sub-tile stride, SBO = 8 rows) descriptor right for a 32B-swizzled MN-major operand whose N extent is
// split across the pair; (3) do the 2-CTA TMA loads of such boxes land where the descriptor expects them.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../../diffsim_b200/csrc -o ubench_pair_pv ubench_pair_pv.cu -lcuda
// run:   ./ubench_pair_pv         (prints the maximum error; expect ~1e-3 relative to |O| ~ 10)
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace ds {
// minimal stand-ins for what ds_ptx.cuh expects from ds_host.h
}
#include "ds_ptx.cuh"

using namespace ds;

constexpr int kM = 256, kK = 128, kN = 160;       // per pair
constexpr int kSubCols = 16, kSubBytes = kSubCols * 2;   // 32-byte swizzle span
constexpr int kSubTileBytes = kK * kSubBytes;             // 128 rows x 32 B = 4 KB
constexpr int kNHalf = kN / 2, kSubTiles = kNHalf / kSubCols;   // 80 columns = 5 sub-tiles per CTA

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_pv_kernel(const __grid_constant__ CUtensorMap map_v, const __half* __restrict__ P, float* __restrict__ O) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSubTiles * kSubTileBytes);
  uint64_t* v_full = bars;        // leader's copy counts both CTAs' bytes
  uint64_t* o_full = bars + 1;    // multicast commit
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    mbar_init(v_full, 1);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc_2cta(tmem_slot, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kTmemP = 0, kTmemO = 256;

  // ---- P: thread = row (TMEM lane) of this CTA's 128 rows; 32 kv columns -> 16 packed words per chunk
  {
    const int row = warp * 32 + lane;
    const __half* prow = P + (size_t)(rank * 128 + row) * kK;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
#pragma unroll
    for (int chunk = 0; chunk < kK / 32; ++chunk) {
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const __half2 h2 = __halves2half2(prow[chunk * 32 + 2 * j], prow[chunk * 32 + 2 * j + 1]);
        pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      tmem_st_x16(tmem_base + lane_addr + kTmemP + chunk * 32, pk);
    }
    tmem_wait_st();
    tc_fence_before_sync();
  }
  // ---- V: this CTA's 80 columns, five 16-column boxes, bytes counted on the leader's barrier
  if (threadIdx.x == 0) {
    if (rank == 0) mbar_arrive_expect_tx(v_full, 2 * kSubTiles * kSubTileBytes);
    for (int s = 0; s < kSubTiles; ++s)
      tma_load_2d_2cta(smem + s * kSubTileBytes, &map_v, v_full, (int)rank * kNHalf + s * kSubCols, 0);
  }
  cluster_sync_all();   // both CTAs' P is in TMEM before the leader issues
  if (rank == 0 && warp == 1 && elect_one()) {
    mbar_wait(v_full, 0);
    tc_fence_after_sync();
    const uint32_t idesc = umma_idesc_f16(0u /*f16*/, kM, kN, 0 /*A K-major (TMEM)*/, 1 /*B MN-major*/);
    // MN-major B: LBO = distance between 16-column sub-tiles, SBO = distance between 8-row groups along K
    const uint64_t v_desc0 = umma_smem_desc(smem_u32(smem), kSubTileBytes, 8 * kSubBytes, UMMA_SW32);
    for (int ks = 0; ks < kK / 16; ++ks) {
      const uint32_t a_tmem = tmem_base + kTmemP + (ks >> 1) * 32 + (ks & 1) * 8;
      // 2-CTA TS form: same instruction as umma_f16_ts with cta_group::2
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
          "}\n" ::"r"(tmem_base + kTmemO),
          "r"(a_tmem), "l"(v_desc0 + (uint64_t)((ks * 16 * kSubBytes) >> 4)), "r"(idesc), "r"(ks > 0 ? 1u : 0u)
          : "memory");
    }
    umma_commit_2cta(o_full);
  }
  mbar_wait(o_full, 0);
  tc_fence_after_sync();
  {
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    float* orow = O + (size_t)(rank * 128 + row) * kN;
#pragma unroll 1
    for (int c = 0; c < kN / 16; ++c) {
      uint32_t v[16];
      tmem_ld_x16(tmem_base + lane_addr + kTmemO + c * 16, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) orow[c * 16 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_2cta(tmem_base, 512);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  std::vector<__half> hP((size_t)kM * kK), hV((size_t)kK * kN);
  std::vector<double> ref((size_t)kM * kN, 0.0);
  srand(7);
  for (auto& x : hP) x = __float2half((float)rand() / RAND_MAX);                 // "probabilities" in [0, 1]
  for (auto& x : hV) x = __float2half(2.0f * rand() / RAND_MAX - 1.0f);
  for (int i = 0; i < kM; ++i)
    for (int k = 0; k < kK; ++k) {
      const double p = __half2float(hP[(size_t)i * kK + k]);
      for (int n = 0; n < kN; ++n) ref[(size_t)i * kN + n] += p * __half2float(hV[(size_t)k * kN + n]);
    }
  __half *dP, *dV;
  float* dO;
  cudaMalloc(&dP, hP.size() * 2);
  cudaMalloc(&dV, hV.size() * 2);
  cudaMalloc(&dO, ref.size() * 4);
  cudaMemcpy(dP, hP.data(), hP.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dV, hV.data(), hV.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0, ref.size() * 4);

  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    printf("no cuTensorMapEncodeTiled\n");
    return 2;
  }
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)kN, (cuuint64_t)kK};      // fastest first: columns (N), rows (K)
  cuuint64_t strides[1] = {(cuuint64_t)kN * 2};
  cuuint32_t box[2] = {(cuuint32_t)kSubCols, (cuuint32_t)kK};
  cuuint32_t es[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeFn>(fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dV, dims, strides, box, es,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return 2;
  }
  const size_t smem_bytes = 1024 + kSubTiles * kSubTileBytes + 256;
  cudaFuncSetAttribute(pair_pv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  pair_pv_kernel<<<2, 128, smem_bytes>>>(map, dP, dO);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("kernel failed: %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> hO(ref.size());
  cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
  double max_err = 0, max_ref = 0;
  int bad_r = -1, bad_c = -1;
  for (int i = 0; i < kM; ++i)
    for (int n = 0; n < kN; ++n) {
      const double d = std::fabs((double)hO[(size_t)i * kN + n] - ref[(size_t)i * kN + n]);
      if (d > max_err) {
        max_err = d;
        bad_r = i;
        bad_c = n;
      }
      max_ref = std::fmax(max_ref, std::fabs(ref[(size_t)i * kN + n]));
    }
  printf("pair PV (cta_group::2, A from TMEM, B N-split with 32B swizzle): max |err| %.3e at (%d, %d), max |ref| %.3f -> %s\n",
         max_err, bad_r, bad_c, max_ref, max_err < 2e-3 * max_ref ? "OK" : "MISMATCH");
  return max_err < 2e-3 * max_ref ? 0 : 1;
}
