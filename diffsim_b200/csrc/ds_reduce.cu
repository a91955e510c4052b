// K2 -- batched similarity reductions (HBM-bound).
//
// One pass over x and y with 128-bit loads, fp32 accumulation, warp-shuffle +
// shared-memory block reduction, and a deterministic "last CTA finishes" second
// stage (fixed summation order, no float atomics).
//
// Replaces F.cosine_similarity / F.mse_loss on flattened tensors
// (diffsim/diffsim.py:182-197) and min_max_normalize + cosine
// (metrics/diffeats.py:136-140,202-205).  Algorithmic bytes per row pair:
// 2 * E * sizeof(dtype) read + 4 written.
#include "ds_host.h"
#include "ds_ptx.cuh"

#include <float.h>

namespace ds {

constexpr int kRedThreads = 256;
constexpr int kRedUnroll = 4;
constexpr int kRedMaxChunks = 64;
constexpr int kRedSlots = 12;  // floats per partial record

template <typename T>
struct Vec16;
template <>
struct Vec16<__half> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void cvt(const uint4& u, float (&f)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __half22float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  static __device__ __forceinline__ float one(const __half* p) { return __half2float(*p); }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void cvt(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  static __device__ __forceinline__ float one(const __nv_bfloat16* p) { return __bfloat162float(*p); }
};
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void cvt(const uint4& u, float (&f)[4]) {
    f[0] = __uint_as_float(u.x);
    f[1] = __uint_as_float(u.y);
    f[2] = __uint_as_float(u.z);
    f[3] = __uint_as_float(u.w);
  }
  static __device__ __forceinline__ float one(const float* p) { return *p; }
};

// streaming 128-bit load: read once, do not pollute L1
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

template <int MODE>
struct Acc {
  // COSINE: dot, xx, yy.  MSE: sq.  MINMAX: dot, xx, yy, sx, sy, mnx, mxx, mny, mxy
  float dot = 0.f, xx = 0.f, yy = 0.f, sx = 0.f, sy = 0.f, sq = 0.f;
  float mnx = FLT_MAX, mxx = -FLT_MAX, mny = FLT_MAX, mxy = -FLT_MAX;
  __device__ __forceinline__ void add(float a, float b) {
    if constexpr (MODE == DS_SIM_MSE) {
      float d = a - b;
      sq = fmaf(d, d, sq);
    } else {
      dot = fmaf(a, b, dot);
      xx = fmaf(a, a, xx);
      yy = fmaf(b, b, yy);
      if constexpr (MODE == DS_SIM_MINMAX_COSINE) {
        sx += a;
        sy += b;
        mnx = fminf(mnx, a);
        mxx = fmaxf(mxx, a);
        mny = fminf(mny, b);
        mxy = fmaxf(mxy, b);
      }
    }
  }
};

// Final similarity from the summed statistics (double: the min-max form cancels).
template <int MODE>
__device__ float finish_stats(const float* s, double E) {
  if constexpr (MODE == DS_SIM_MSE) {
    return (float)((double)s[5] / E);
  } else if constexpr (MODE == DS_SIM_COSINE) {
    double nx = fmax(sqrt((double)s[1]), 1e-8), ny = fmax(sqrt((double)s[2]), 1e-8);
    return (float)((double)s[0] / (nx * ny));
  } else {
    // x' = (x - ax)/cx, y' = (y - ay)/cy   (metrics/diffeats.py:136-140)
    double ax = s[6], cx = (double)s[7] - (double)s[6];
    double ay = s[8], cy = (double)s[9] - (double)s[8];
    double dot = ((double)s[0] - ay * (double)s[3] - ax * (double)s[4] + E * ax * ay) / (cx * cy);
    double xx = ((double)s[1] - 2.0 * ax * (double)s[3] + E * ax * ax) / (cx * cx);
    double yy = ((double)s[2] - 2.0 * ay * (double)s[4] + E * ay * ay) / (cy * cy);
    double nx = fmax(sqrt(fmax(xx, 0.0)), 1e-8), ny = fmax(sqrt(fmax(yy, 0.0)), 1e-8);
    return (float)(dot / (nx * ny));
  }
}

template <typename T, int MODE>
__global__ void __launch_bounds__(kRedThreads)
pair_reduce_kernel(const T* __restrict__ x, const T* __restrict__ y, int64_t E, int64_t xs, int64_t ys,
                   int chunks, int64_t chunk_elems, int vec_ok, float* __restrict__ partials,
                   unsigned int* __restrict__ counters, float* __restrict__ out) {
  constexpr int VN = Vec16<T>::N;
  const int64_t pair = blockIdx.x / chunks;
  const int chunk = blockIdx.x % chunks;
  const T* xp = x + pair * xs;
  const T* yp = y + pair * ys;
  const int64_t e0 = (int64_t)chunk * chunk_elems;
  const int64_t e1 = min(E, e0 + chunk_elems);

  Acc<MODE> acc;
  // vector body: chunk_elems is a multiple of VN * threads * unroll, so e0 keeps the rows' 16-byte alignment;
  // rows that are not 16-byte aligned (vec_ok == 0) take the element-wise loop below for everything
  const int64_t nvec = (vec_ok && e1 > e0) ? (e1 - e0) / VN : 0;
  const uint4* xv = reinterpret_cast<const uint4*>(xp + e0);
  const uint4* yv = reinterpret_cast<const uint4*>(yp + e0);
  int64_t i = threadIdx.x;
  for (; i + (kRedUnroll - 1) * kRedThreads < nvec; i += kRedUnroll * kRedThreads) {
    uint4 a[kRedUnroll], b[kRedUnroll];
#pragma unroll
    for (int u = 0; u < kRedUnroll; ++u) {
      a[u] = ld_stream(xv + i + u * kRedThreads);
      b[u] = ld_stream(yv + i + u * kRedThreads);
    }
#pragma unroll
    for (int u = 0; u < kRedUnroll; ++u) {
      float fa[VN], fb[VN];
      Vec16<T>::cvt(a[u], fa);
      Vec16<T>::cvt(b[u], fb);
#pragma unroll
      for (int k = 0; k < VN; ++k) acc.add(fa[k], fb[k]);
    }
  }
  for (; i < nvec; i += kRedThreads) {
    float fa[VN], fb[VN];
    Vec16<T>::cvt(ld_stream(xv + i), fa);
    Vec16<T>::cvt(ld_stream(yv + i), fb);
#pragma unroll
    for (int k = 0; k < VN; ++k) acc.add(fa[k], fb[k]);
  }
  // scalar tail (only the last chunk can have one)
  for (int64_t e = e0 + nvec * VN + threadIdx.x; e < e1; e += kRedThreads)
    acc.add(Vec16<T>::one(xp + e), Vec16<T>::one(yp + e));

  // block reduction: shuffle inside the warp, shared memory across warps
  __shared__ float sred[kRedThreads / 32][kRedSlots];
  __shared__ bool is_last;
  float v[10] = {acc.dot, acc.xx, acc.yy, acc.sx, acc.sy, acc.sq, acc.mnx, acc.mxx, acc.mny, acc.mxy};
#pragma unroll
  for (int k = 0; k < 6; ++k) v[k] = warp_sum(v[k]);
  if constexpr (MODE == DS_SIM_MINMAX_COSINE) {
    v[6] = warp_min(v[6]);
    v[7] = warp_max(v[7]);
    v[8] = warp_min(v[8]);
    v[9] = warp_max(v[9]);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 10; ++k) sred[warp][k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float r[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) r[k] = sred[0][k];
    for (int w = 1; w < kRedThreads / 32; ++w) {
#pragma unroll
      for (int k = 0; k < 6; ++k) r[k] += sred[w][k];
      r[6] = fminf(r[6], sred[w][6]);
      r[7] = fmaxf(r[7], sred[w][7]);
      r[8] = fminf(r[8], sred[w][8]);
      r[9] = fmaxf(r[9], sred[w][9]);
    }
    float* dst = partials + ((size_t)pair * chunks + chunk) * kRedSlots;
#pragma unroll
    for (int k = 0; k < 10; ++k) dst[k] = r[k];
    __threadfence();
    unsigned int prev = atomicAdd(&counters[pair], 1u);
    is_last = (prev == (unsigned int)(chunks - 1));
  }
  __syncthreads();
  if (!is_last) return;
  // last CTA of this pair: sum the chunk partials in chunk order (deterministic)
  if (threadIdx.x == 0) {
    __threadfence();
    const volatile float* src = partials + (size_t)pair * chunks * kRedSlots;
    float r[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) r[k] = src[k];
    for (int c = 1; c < chunks; ++c) {
      const volatile float* s = src + (size_t)c * kRedSlots;
#pragma unroll
      for (int k = 0; k < 6; ++k) r[k] += s[k];
      r[6] = fminf(r[6], s[6]);
      r[7] = fmaxf(r[7], s[7]);
      r[8] = fminf(r[8], s[8]);
      r[9] = fmaxf(r[9], s[9]);
    }
    out[pair] = finish_stats<MODE>(r, (double)E);
    counters[pair] = 0;  // leave the workspace reusable
  }
}

static void reduce_plan(int64_t n_pairs, int64_t E, int vn, int* chunks, int64_t* chunk_elems) {
  const int64_t quantum = (int64_t)vn * kRedThreads * kRedUnroll;  // elements one CTA sweep covers
  int64_t max_chunks = (E + quantum - 1) / quantum;
  if (max_chunks < 1) max_chunks = 1;
  int64_t want = (2 * 148 + n_pairs - 1) / n_pairs;        // at least two CTAs per SM overall
  int64_t by_size = (E + 131071) / 131072;                 // at most 128 Ki elements per CTA
  int64_t c = want > by_size ? want : by_size;
  if (c > max_chunks) c = max_chunks;
  if (c > kRedMaxChunks) c = kRedMaxChunks;
  if (c < 1) c = 1;
  int64_t ce = (E + c - 1) / c;
  ce = (ce + quantum - 1) / quantum * quantum;
  c = (E + ce - 1) / ce;
  if (c < 1) c = 1;
  *chunks = (int)c;
  *chunk_elems = ce;
}

template <typename T>
static int launch_reduce(const void* x, const void* y, int64_t n_pairs, int64_t E, int64_t xs, int64_t ys,
                         int mode, float* out, float* partials, unsigned int* counters, int chunks,
                         int64_t chunk_elems, int vec_ok, cudaStream_t st) {
  const T* xp = static_cast<const T*>(x);
  const T* yp = static_cast<const T*>(y);
  int64_t blocks = n_pairs * chunks;
  if (blocks > 0x7fffffffLL) return fail(DS_ERR_INVALID, "ds_pair_reduce: too many pairs");
  dim3 grid((unsigned)blocks), block(kRedThreads);
  switch (mode) {
    case DS_SIM_COSINE:
      pair_reduce_kernel<T, DS_SIM_COSINE><<<grid, block, 0, st>>>(xp, yp, E, xs, ys, chunks, chunk_elems, vec_ok, partials,
                                                                   counters, out);
      break;
    case DS_SIM_MSE:
      pair_reduce_kernel<T, DS_SIM_MSE><<<grid, block, 0, st>>>(xp, yp, E, xs, ys, chunks, chunk_elems, vec_ok, partials,
                                                                counters, out);
      break;
    case DS_SIM_MINMAX_COSINE:
      pair_reduce_kernel<T, DS_SIM_MINMAX_COSINE><<<grid, block, 0, st>>>(xp, yp, E, xs, ys, chunks, chunk_elems,
                                                                          vec_ok, partials, counters, out);
      break;
    default:
      return fail(DS_ERR_INVALID, "ds_pair_reduce: bad mode %d", mode);
  }
  DS_CUDA_TRY(cudaGetLastError());
  return DS_OK;
}

}  // namespace ds

extern "C" {

size_t ds_pair_reduce_workspace_bytes(int64_t n_pairs, int64_t E) {
  if (n_pairs <= 0 || E <= 0) return 256;
  // sized for the widest split so that the plan never needs more
  size_t part = (size_t)n_pairs * ds::kRedMaxChunks * ds::kRedSlots * sizeof(float);
  size_t cnt = (size_t)n_pairs * sizeof(unsigned int);
  return ds::align_up(part, 256) + ds::align_up(cnt, 256) + 256;
}

int ds_pair_reduce(const void* x, const void* y, int64_t n_pairs, int64_t E, int64_t x_stride, int64_t y_stride,
                   int dtype, int mode, float* out, void* ws, size_t ws_bytes, void* stream) {
  using namespace ds;
  if (n_pairs < 0 || E <= 0) return fail(DS_ERR_INVALID, "ds_pair_reduce: n_pairs %lld E %lld", (long long)n_pairs, (long long)E);
  if (n_pairs == 0) return DS_OK;
  if (!x || !y || !out) return fail(DS_ERR_INVALID, "ds_pair_reduce: null pointer");
  if (dtype != DS_F16 && dtype != DS_BF16 && dtype != DS_F32) return fail(DS_ERR_INVALID, "ds_pair_reduce: bad dtype %d", dtype);
  if (mode != DS_SIM_COSINE && mode != DS_SIM_MSE && mode != DS_SIM_MINMAX_COSINE)
    return fail(DS_ERR_INVALID, "ds_pair_reduce: bad mode %d", mode);
  const size_t es = elem_size(dtype);
  if (((uintptr_t)x % es) || ((uintptr_t)y % es)) return fail(DS_ERR_INVALID, "ds_pair_reduce: x and y must be element-aligned");
  if (x_stride < E || y_stride < E) return fail(DS_ERR_INVALID, "ds_pair_reduce: row stride smaller than E");
  // 128-bit loads need every row to start on a 16-byte boundary; otherwise the kernel reads element-wise
  const int vec_ok = !(((uintptr_t)x & 15) || ((uintptr_t)y & 15) ||
                       (n_pairs > 1 && ((((size_t)x_stride * es) & 15) || (((size_t)y_stride * es) & 15))));
  int rc = ds_device_ok();
  if (rc != DS_OK) return rc;

  int chunks;
  int64_t chunk_elems;
  reduce_plan(n_pairs, E, dtype == DS_F32 ? 4 : 8, &chunks, &chunk_elems);
  Workspace w(ws, ws_bytes);
  float* partials = static_cast<float*>(w.take((size_t)n_pairs * chunks * kRedSlots * sizeof(float)));
  unsigned int* counters = static_cast<unsigned int*>(w.take((size_t)n_pairs * sizeof(unsigned int)));
  if (!partials || !counters)
    return fail(DS_ERR_WORKSPACE, "ds_pair_reduce: workspace too small (%zu bytes given, need %zu)", ws_bytes,
                ds_pair_reduce_workspace_bytes(n_pairs, E));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DS_CUDA_TRY(cudaMemsetAsync(counters, 0, (size_t)n_pairs * sizeof(unsigned int), st));
  if (dtype == DS_F16)
    return launch_reduce<__half>(x, y, n_pairs, E, x_stride, y_stride, mode, out, partials, counters, chunks, chunk_elems, vec_ok, st);
  if (dtype == DS_BF16)
    return launch_reduce<__nv_bfloat16>(x, y, n_pairs, E, x_stride, y_stride, mode, out, partials, counters, chunks, chunk_elems, vec_ok, st);
  return launch_reduce<float>(x, y, n_pairs, E, x_stride, y_stride, mode, out, partials, counters, chunks, chunk_elems, vec_ok, st);
}

}  // extern "C"
