// f1: the parameter-free cross-attention of Pretrain.local_text_token_alignment_loss,
// models/model_pretrain_finetune_v0520.py:509-511 - every text token attends over the patch tokens of its own
// sample, att = softmax(T V^T / sqrt(D)), O = att V - forward and backward.  The token-level InfoNCE that follows
// (:514-525) is, per sample, the G loss with identity ids, and runs on the batched small-path kernels.
//
// Reference sizes are tiny (B = 32, L ~ 99 text tokens, P = 49 patches, D = 768: ~1 GFLOP per step), so these are
// fp32 SIMT kernels, one CTA per (8 tokens, sample) resp. (8 patches, sample) so that every patch / token row
// read from L2 serves 8 outputs; latency-bound; everything is deterministic (no atomics).
// Layouts: text [B, L, D], image [B, P, D], att / ds [B, L, P], out [B, L, D], all contiguous fp32.
#include "evk_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxP = 1024;                    // patches per sample held in shared memory

constexpr int kTile = 8;                       // tokens (or patches) per CTA: each patch / token row read is shared by 8

__global__ void __launch_bounds__(kThreads)
local_attend_fwd_kernel(const float* __restrict__ text, const float* __restrict__ image, int l_tokens, int p_tokens, int d,
                        float inv_sqrt_d, float* __restrict__ att, float* __restrict__ out) {
  extern __shared__ float sm[];
  float* t_s = sm;                              // [kTile][d]
  float* sc = sm + (size_t)kTile * d;           // [kTile][p_tokens]
  const int l0 = blockIdx.x * kTile, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* v = image + (int64_t)b * p_tokens * d;
  for (int i = threadIdx.x; i < kTile * d; i += kThreads) {
    const int k = i / d, c = i - k * d;
    t_s[i] = (l0 + k < l_tokens) ? text[((int64_t)b * l_tokens + l0 + k) * d + c] : 0.f;
  }
  __syncthreads();
  for (int p = warp; p < p_tokens; p += kWarps) {                // scores of the 8 tokens against patch p
    float acc[kTile];
#pragma unroll
    for (int k = 0; k < kTile; ++k) acc[k] = 0.f;
    const float* vr = v + (int64_t)p * d;
    for (int c = lane; c < d; c += 32) {
      const float vv = __ldg(vr + c);
#pragma unroll
      for (int k = 0; k < kTile; ++k) acc[k] = fmaf(t_s[k * d + c], vv, acc[k]);
    }
#pragma unroll
    for (int k = 0; k < kTile; ++k) {
      const float s = warp_sum(acc[k]);
      if (lane == 0) sc[k * p_tokens + p] = s * inv_sqrt_d;
    }
  }
  __syncthreads();
  {                                                                // softmax over the patches: warp k <-> token k
    float* row = sc + warp * p_tokens;
    float m = -INFINITY;
    for (int p = lane; p < p_tokens; p += 32) m = fmaxf(m, row[p]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float z = 0.f;
    for (int p = lane; p < p_tokens; p += 32) {
      const float e = expf(row[p] - m);
      row[p] = e;
      z += e;
    }
    z = warp_sum(z);
    const float inv_z = 1.f / z;
    const bool ok = l0 + warp < l_tokens;
    float* a_row = att + ((int64_t)b * l_tokens + l0 + warp) * p_tokens;
    for (int p = lane; p < p_tokens; p += 32) {
      const float a = row[p] * inv_z;
      row[p] = a;
      if (ok) a_row[p] = a;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += kThreads) {              // out rows of the 8 tokens
    float acc[kTile];
#pragma unroll
    for (int k = 0; k < kTile; ++k) acc[k] = 0.f;
    for (int p = 0; p < p_tokens; ++p) {
      const float vv = __ldg(v + (int64_t)p * d + c);
#pragma unroll
      for (int k = 0; k < kTile; ++k) acc[k] = fmaf(sc[k * p_tokens + p], vv, acc[k]);
    }
#pragma unroll
    for (int k = 0; k < kTile; ++k)
      if (l0 + k < l_tokens) out[((int64_t)b * l_tokens + l0 + k) * d + c] = acc[k];
  }
}

// per (8 tokens, sample): ds = att * (dA - sum(dA * att)) / sqrt(D), dA_p = dO . V_p;  d_text[l] += ds V
__global__ void __launch_bounds__(kThreads)
local_attend_bwd_token_kernel(const float* __restrict__ image, const float* __restrict__ att,
                              const float* __restrict__ d_out, int l_tokens, int p_tokens, int d, float inv_sqrt_d,
                              float* __restrict__ ds, float* __restrict__ d_text) {
  extern __shared__ float sm[];
  float* g_s = sm;                              // [kTile][d]  dO of the 8 tokens
  float* da = sm + (size_t)kTile * d;           // [kTile][p_tokens]
  const int l0 = blockIdx.x * kTile, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* v = image + (int64_t)b * p_tokens * d;
  for (int i = threadIdx.x; i < kTile * d; i += kThreads) {
    const int k = i / d, c = i - k * d;
    g_s[i] = (l0 + k < l_tokens) ? d_out[((int64_t)b * l_tokens + l0 + k) * d + c] : 0.f;
  }
  __syncthreads();
  for (int p = warp; p < p_tokens; p += kWarps) {
    float acc[kTile];
#pragma unroll
    for (int k = 0; k < kTile; ++k) acc[k] = 0.f;
    const float* vr = v + (int64_t)p * d;
    for (int c = lane; c < d; c += 32) {
      const float vv = __ldg(vr + c);
#pragma unroll
      for (int k = 0; k < kTile; ++k) acc[k] = fmaf(g_s[k * d + c], vv, acc[k]);
    }
#pragma unroll
    for (int k = 0; k < kTile; ++k) {
      const float s = warp_sum(acc[k]);
      if (lane == 0) da[k * p_tokens + p] = s;
    }
  }
  __syncthreads();
  {                                                                // warp k <-> token k
    const bool ok = l0 + warp < l_tokens;
    const int64_t tok = (int64_t)b * l_tokens + l0 + warp;
    float* row = da + warp * p_tokens;
    float dot = 0.f;
    for (int p = lane; p < p_tokens; p += 32) dot = fmaf(row[p], ok ? att[tok * p_tokens + p] : 0.f, dot);
    dot = warp_sum(dot);
    for (int p = lane; p < p_tokens; p += 32) {
      const float v_ds = ok ? att[tok * p_tokens + p] * (row[p] - dot) * inv_sqrt_d : 0.f;
      row[p] = v_ds;
      if (ok) ds[tok * p_tokens + p] = v_ds;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += kThreads) {
    float acc[kTile];
#pragma unroll
    for (int k = 0; k < kTile; ++k) acc[k] = 0.f;
    for (int p = 0; p < p_tokens; ++p) {
      const float vv = __ldg(v + (int64_t)p * d + c);
#pragma unroll
      for (int k = 0; k < kTile; ++k) acc[k] = fmaf(da[k * p_tokens + p], vv, acc[k]);
    }
#pragma unroll
    for (int k = 0; k < kTile; ++k)
      if (l0 + k < l_tokens) d_text[((int64_t)b * l_tokens + l0 + k) * d + c] += acc[k];
  }
}

// per (8 patches, sample): d_image[p] = sum_l ( att[l, p] dO[l] + ds[l, p] T[l] )   (fixed order over l)
__global__ void __launch_bounds__(kThreads)
local_attend_bwd_patch_kernel(const float* __restrict__ text, const float* __restrict__ att, const float* __restrict__ ds,
                              const float* __restrict__ d_out, int l_tokens, int p_tokens, int d,
                              float* __restrict__ d_image) {
  extern __shared__ float sm[];
  float* a_s = sm;                              // [l_tokens][kTile] att[:, p0..p0+7]
  float* s_s = sm + (size_t)l_tokens * kTile;   // [l_tokens][kTile] ds[:, p0..p0+7]
  const int p0 = blockIdx.x * kTile, b = blockIdx.y;
  for (int i = threadIdx.x; i < l_tokens * kTile; i += kThreads) {
    const int l = i / kTile, k = i - l * kTile;
    const bool ok = p0 + k < p_tokens;
    const int64_t j = ((int64_t)b * l_tokens + l) * p_tokens + p0 + k;
    a_s[i] = ok ? att[j] : 0.f;
    s_s[i] = ok ? ds[j] : 0.f;
  }
  __syncthreads();
  const float* t = text + (int64_t)b * l_tokens * d;
  const float* g = d_out + (int64_t)b * l_tokens * d;
  for (int c = threadIdx.x; c < d; c += kThreads) {
    float acc[kTile];
#pragma unroll
    for (int k = 0; k < kTile; ++k) acc[k] = 0.f;
    for (int l = 0; l < l_tokens; ++l) {
      const float gv = __ldg(g + (int64_t)l * d + c), tv = __ldg(t + (int64_t)l * d + c);
#pragma unroll
      for (int k = 0; k < kTile; ++k) acc[k] = fmaf(a_s[l * kTile + k], gv, fmaf(s_s[l * kTile + k], tv, acc[k]));
    }
#pragma unroll
    for (int k = 0; k < kTile; ++k)
      if (p0 + k < p_tokens) d_image[((int64_t)b * p_tokens + p0 + k) * d + c] = acc[k];
  }
}

int check_shape(int64_t batch, int64_t l, int64_t p, int64_t d) {
  EVK_REQUIRE(batch >= 1 && batch <= 65535 && l >= 1 && p >= 1 && d >= 1, "evk_local_attend: empty shape or batch > 65535");
  EVK_REQUIRE(p <= kMaxP && l <= 4096 && d <= 4096, "evk_local_attend: supports p <= %d patches, l <= 4096 tokens, d <= 4096", kMaxP);
  static_assert(kWarps == kTile, "one warp per token of the tile");
  return EVK_OK;
}

}  // namespace

extern "C" int evk_local_attend_fwd(const float* text, const float* image, int64_t batch, int64_t l, int64_t p, int64_t d,
                                    float* att, float* out, evk_stream_t stream) {
  EVK_REQUIRE(text && image && att && out, "evk_local_attend_fwd: null pointer");
  int rc = check_shape(batch, l, p, d);
  if (rc != EVK_OK) return rc;
  const size_t smem = sizeof(float) * (size_t)kTile * (size_t)(d + p);
  EVK_REQUIRE(smem <= 200 * 1024, "evk_local_attend_fwd: d + p too large for shared memory");
  EVK_CUDA(cudaFuncSetAttribute(local_attend_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  local_attend_fwd_kernel<<<dim3((unsigned)((l + kTile - 1) / kTile), (unsigned)batch), kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
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
  const size_t smem1 = sizeof(float) * (size_t)kTile * (size_t)(d + p);
  EVK_REQUIRE(smem1 <= 200 * 1024, "evk_local_attend_bwd: d + p too large for shared memory");
  EVK_CUDA(cudaFuncSetAttribute(local_attend_bwd_token_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  local_attend_bwd_token_kernel<<<dim3((unsigned)((l + kTile - 1) / kTile), (unsigned)batch), kThreads, smem1, s>>>(
      image, att, d_out, (int)l, (int)p, (int)d, 1.f / sqrtf((float)d), ds, d_text);
  EVK_CHECK_LAUNCH("local_attend_bwd_token");
  const size_t smem2 = sizeof(float) * (size_t)(2 * l * kTile);
  EVK_CUDA(cudaFuncSetAttribute(local_attend_bwd_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  local_attend_bwd_patch_kernel<<<dim3((unsigned)((p + kTile - 1) / kTile), (unsigned)batch), kThreads, smem2, s>>>(text, att, ds, d_out, (int)l, (int)p,
                                                                                          (int)d, d_image);
  EVK_CHECK_LAUNCH("local_attend_bwd_patch");
  return EVK_OK;
}
