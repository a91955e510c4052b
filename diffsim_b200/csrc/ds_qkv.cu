// K4 -- QKV projection of the hooked self-attention layer (tensor-core bound).
//
//   [q | k | v][r, :] = hidden[r, :] . W^T (+ bias)          W = [W_q; W_k; W_v]  (nn.Linear layout: [out, in])
//
// The capture step of the reference: attn.to_q / to_k / to_v on the hook's input (diffsim/hacked_attn.py:61-69;
// the head split of :74-77 is a view and costs nothing here: the outputs ARE the (B,S,H*D) memory the attention
// kernel reads through (B,H,S,D) strides) and DiT's fused module.qkv(x) (diffsim/diffsim_dit.py:21-23, packed
// (B,N,3*H*D) output).  With it the drop-in boundary moves from Q/K/V to the hook INPUT: a third of the bytes.
// One pass of the persistent tcgen05 GEMM core (ds_gemm.cuh), fp32 accumulation, bias added in fp32, rounded once
// to the input dtype (as a cuBLAS-backed nn.Linear does).  Algorithmic work: 2 * rows * C_in * n_out flops.
#include "ds_gemm.cuh"

extern "C" {

int ds_qkv_project(const void* hidden, int64_t n_rows, int64_t ld_hidden, int64_t c_in, const void* weight,
                   int64_t ld_weight, const void* bias, int64_t n_out, int64_t cols_per_out, void* const* out,
                   const int64_t* ld_out, int dtype, void* stream) {
  using namespace ds;
  if (n_rows < 0 || c_in <= 0 || n_out <= 0 || cols_per_out <= 0) return fail(DS_ERR_INVALID, "ds_qkv_project: bad sizes");
  if (n_rows == 0) return DS_OK;
  if (!hidden || !weight || !out || !ld_out) return fail(DS_ERR_INVALID, "ds_qkv_project: null pointer");
  if (dtype != DS_F16 && dtype != DS_BF16) return fail(DS_ERR_UNSUPPORTED, "ds_qkv_project: dtype must be f16 or bf16");
  if (n_out % cols_per_out) return fail(DS_ERR_INVALID, "ds_qkv_project: n_out must be a multiple of cols_per_out");
  const int64_t n_t = n_out / cols_per_out;
  if (n_t < 1 || n_t > 3) return fail(DS_ERR_INVALID, "ds_qkv_project: 1 to 3 output tensors (got %lld)", (long long)n_t);
  if ((cols_per_out & 7) || (c_in & 7) || (ld_hidden & 7) || (ld_weight & 7) || ld_hidden < c_in || ld_weight < c_in)
    return fail(DS_ERR_INVALID, "ds_qkv_project: channel counts and leading dimensions must be multiples of 8 elements");
  if (((uintptr_t)hidden & 15) || ((uintptr_t)weight & 15) || (bias && ((uintptr_t)bias & 15)))
    return fail(DS_ERR_INVALID, "ds_qkv_project: pointers must be 16-byte aligned");
  if (n_rows > INT32_MAX || n_out > INT32_MAX) return fail(DS_ERR_INVALID, "ds_qkv_project: sizes too large");
  GemmParams gp = {};
  for (int64_t t = 0; t < n_t; ++t) {
    if (!out[t] || ((uintptr_t)out[t] & 15) || (ld_out[t] & 7) || ld_out[t] < cols_per_out)
      return fail(DS_ERR_INVALID, "ds_qkv_project: output %lld needs a 16-byte aligned pointer and ld >= cols_per_out, multiple of 8",
                  (long long)t);
    gp.out[t] = out[t];
    gp.ld_out[t] = ld_out[t];
  }
  int rc = ds_device_ok();
  if (rc != DS_OK) return rc;
  gp.splits = 1;
  gp.use_pair = -1;
  gp.cols_per_out = (int)cols_per_out;
  gp.bias = bias;
  return launch_gemm_tn<GEMM_EPI_16>(hidden, n_rows, ld_hidden, weight, n_out, ld_weight, c_in, dtype, gp,
                                     static_cast<cudaStream_t>(stream));
}

}  // extern "C"
