// f1: the parameter-free cross-attention of Pretrain.local_text_token_alignment_loss,
// models/model_pretrain_finetune_v0520.py:509-511 - every text token attends over the patch tokens of its own
// sample, att = softmax(T V^T / sqrt(D)), O = att V - forward and backward.  The token-level InfoNCE that follows
// (:514-525) is, per sample, the G loss with identity ids, and runs on the batched small-path kernels.
//
// Reference sizes are tiny (B = 32, L ~ 99 text tokens, P = 49 patches, D = 768: ~1 GFLOP per step), so these are
// fp32 SIMT kernels, one CTA per (token, sample), latency-bound; everything is deterministic (no atomics).
// Layouts: text [B, L, D], image [B, P, D], att / ds [B, L, P], out [B, L, D], all contiguous fp32.
#include "evk_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxP = 1024;                    // patches per sample held in shared memory

// dot(x_s[0..d), row[0..d)) with the whole warp; x_s in shared memory
__device__ __forceinline__ float warp_dot(const float* __restrict__ x_s, const float* __restrict__ row, int d, int lane) {
  float acc = 0.f;
  for (int c = lane; c < d; c += 32) acc = fmaf(x_s[c], __ldg(row + c), acc);
  return warp_sum(acc);
}

__global__ void __launch_bounds__(kThreads)
local_attend_fwd_kernel(const float* __restrict__ text, const float* __restrict__ image, int l_tokens, int p_tokens, int d,
                        float inv_sqrt_d, float* __restrict__ att, float* __restrict__ out) {
  extern __shared__ float sm[];
  float* t_s = sm;                 // [d]
  float* sc = sm + d;              // [p_tokens]
  __shared__ float s_red[2];
  const int l = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* t_row = text + ((int64_t)b * l_tokens + l) * d;
  const float* v = image + (int64_t)b * p_tokens * d;
  for (int c = threadIdx.x; c < d; c += kThreads) t_s[c] = t_row[c];
  __syncthreads();
  for (int p = warp; p < p_tokens; p += kWarps) {
    const float s = warp_dot(t_s, v + (int64_t)p * d, d, lane);
    if (lane == 0) sc[p] = s * inv_sqrt_d;
  }
  __syncthreads();
  if (warp == 0) {                 // softmax over the patches (F.softmax: max-subtracted)
    float m = -INFINITY;
    for (int p = lane; p < p_tokens; p += 32) m = fmaxf(m, sc[p]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float z = 0.f;
    for (int p = lane; p < p_tokens; p += 32) {
      const float e = expf(sc[p] - m);
      sc[p] = e;
      z += e;
    }
    z = warp_sum(z);
    if (lane == 0) s_red[0] = 1.f / z;
  }
  __syncthreads();
  const float inv_z = s_red[0];
  float* a_row = att + ((int64_t)b * l_tokens + l) * p_tokens;
  for (int p = threadIdx.x; p < p_tokens; p += kThreads) {
    const float a = sc[p] * inv_z;
    sc[p] = a;
    a_row[p] = a;
  }
  __syncthreads();
  float* o_row = out + ((int64_t)b * l_tokens + l) * d;
  for (int c = threadIdx.x; c < d; c += kThreads) {
    float acc = 0.f;
    for (int p = 0; p < p_tokens; ++p) acc = fmaf(sc[p], __ldg(v + (int64_t)p * d + c), acc);
    o_row[c] = acc;
  }
}

// per (token, sample): ds = att * (dA - sum(dA * att)) / sqrt(D), dA_p = dO . V_p;  d_text[l] += ds V
__global__ void __launch_bounds__(kThreads)
local_attend_bwd_token_kernel(const float* __restrict__ image, const float* __restrict__ att,
                              const float* __restrict__ d_out, int l_tokens, int p_tokens, int d, float inv_sqrt_d,
                              float* __restrict__ ds, float* __restrict__ d_text) {
  extern __shared__ float sm[];
  float* g_s = sm;                 // [d]  dO of this token
  float* da = sm + d;              // [p_tokens]
  __shared__ float s_dot;
  const int l = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t tok = (int64_t)b * l_tokens + l;
  const float* v = image + (int64_t)b * p_tokens * d;
  const float* a_row = att + tok * p_tokens;
  for (int c = threadIdx.x; c < d; c += kThreads) g_s[c] = d_out[tok * d + c];
  __syncthreads();
  for (int p = warp; p < p_tokens; p += kWarps) {
    const float s = warp_dot(g_s, v + (int64_t)p * d, d, lane);
    if (lane == 0) da[p] = s;
  }
  __syncthreads();
  if (warp == 0) {
    float acc = 0.f;
    for (int p = lane; p < p_tokens; p += 32) acc = fmaf(da[p], a_row[p], acc);
    acc = warp_sum(acc);
    if (lane == 0) s_dot = acc;
  }
  __syncthreads();
  const float dot = s_dot;
  for (int p = threadIdx.x; p < p_tokens; p += kThreads) {
    const float v_ds = a_row[p] * (da[p] - dot) * inv_sqrt_d;
    da[p] = v_ds;
    ds[tok * p_tokens + p] = v_ds;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += kThreads) {
    float acc = 0.f;
    for (int p = 0; p < p_tokens; ++p) acc = fmaf(da[p], __ldg(v + (int64_t)p * d + c), acc);
    d_text[tok * d + c] += acc;
  }
}

// per (patch, sample): d_image[p] = sum_l ( att[l, p] dO[l] + ds[l, p] T[l] )   (fixed order over l)
__global__ void __launch_bounds__(kThreads)
local_attend_bwd_patch_kernel(const float* __restrict__ text, const float* __restrict__ att, const float* __restrict__ ds,
                              const float* __restrict__ d_out, int l_tokens, int p_tokens, int d,
                              float* __restrict__ d_image) {
  extern __shared__ float sm[];
  float* a_s = sm;                 // [l_tokens] att[:, p]
  float* s_s = sm + l_tokens;      // [l_tokens] ds[:, p]
  const int p = blockIdx.x, b = blockIdx.y;
  for (int l = threadIdx.x; l < l_tokens; l += kThreads) {
    const int64_t i = ((int64_t)b * l_tokens + l) * p_tokens + p;
    a_s[l] = att[i];
    s_s[l] = ds[i];
  }
  __syncthreads();
  const float* t = text + (int64_t)b * l_tokens * d;
  const float* g = d_out + (int64_t)b * l_tokens * d;
  for (int c = threadIdx.x; c < d; c += kThreads) {
    float acc = 0.f;
    for (int l = 0; l < l_tokens; ++l) {
      acc = fmaf(a_s[l], __ldg(g + (int64_t)l * d + c), acc);
      acc = fmaf(s_s[l], __ldg(t + (int64_t)l * d + c), acc);
    }
    d_image[((int64_t)b * p_tokens + p) * d + c] = acc;
  }
}

int check_shape(int64_t batch, int64_t l, int64_t p, int64_t d) {
  EVK_REQUIRE(batch >= 1 && batch <= 65535 && l >= 1 && p >= 1 && d >= 1, "evk_local_attend: empty shape or batch > 65535");
  EVK_REQUIRE(p <= kMaxP && l <= 4096 && d <= 8192, "evk_local_attend: supports p <= %d patches, l <= 4096 tokens, d <= 8192", kMaxP);
  return EVK_OK;
}

}  // namespace

extern "C" int evk_local_attend_fwd(const float* text, const float* image, int64_t batch, int64_t l, int64_t p, int64_t d,
                                    float* att, float* out, evk_stream_t stream) {
  EVK_REQUIRE(text && image && att && out, "evk_local_attend_fwd: null pointer");
  int rc = check_shape(batch, l, p, d);
  if (rc != EVK_OK) return rc;
  const size_t smem = sizeof(float) * (size_t)(d + p);
  EVK_CUDA(cudaFuncSetAttribute(local_attend_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  local_attend_fwd_kernel<<<dim3((unsigned)l, (unsigned)batch), kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      text, image, (int)l, (int)p, (int)d, 1.f / sqrtf((float)d), att, out);
  EVK_CHECK_LAUNCH("local_attend_fwd");
  return EVK_OK;
}

extern "C" int evk_local_attend_bwd(const float* text, const float* image, const float* att, const float* d_out,
                                    int64_t batch, int64_t l, int64_t p, int64_t d, float* ds, float* d_text,
                                    float* d_image, evk_stream_t stream) {
  EVK_REQUIRE(text && image && att && d_out && ds && d_text && d_image, "evk_local_attend_bwd: null pointer");
  int rc = check_shape(batch, l, p, d);
  if (rc != EVK_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem1 = sizeof(float) * (size_t)(d + p);
  EVK_CUDA(cudaFuncSetAttribute(local_attend_bwd_token_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  local_attend_bwd_token_kernel<<<dim3((unsigned)l, (unsigned)batch), kThreads, smem1, s>>>(
      image, att, d_out, (int)l, (int)p, (int)d, 1.f / sqrtf((float)d), ds, d_text);
  EVK_CHECK_LAUNCH("local_attend_bwd_token");
  const size_t smem2 = sizeof(float) * (size_t)(2 * l);
  EVK_CUDA(cudaFuncSetAttribute(local_attend_bwd_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  local_attend_bwd_patch_kernel<<<dim3((unsigned)p, (unsigned)batch), kThreads, smem2, s>>>(text, att, ds, d_out, (int)l, (int)p,
                                                                                          (int)d, d_image);
  EVK_CHECK_LAUNCH("local_attend_bwd_patch");
  return EVK_OK;
}
