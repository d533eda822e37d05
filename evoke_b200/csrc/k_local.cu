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

// ---- per-sample token-level InfoNCE (:518-525) for L <= 128: register-blocked fp32 kernels, one sample per CTA (row) ----
constexpr int kTS = 128;                       // tokens per sample held by one CTA
constexpr int kKC = 16;                        // feature chunk of the similarity product

// E[b, i, j] = exp(That_i . Ohat_j / tau - 1/tau), its row sums, column sums and the diagonal logits.
// 256 threads = 16 x 16; thread (ty, tx) owns rows 8 ty .. 8 ty + 7 and columns tx, tx + 16, .. (interleaved: the
// shared-memory reads of the column operand and the stores of E are then conflict-free / coalesced).
__global__ void __launch_bounds__(kThreads)
token_sim_fwd_kernel(const float* __restrict__ th, const float* __restrict__ oh, int l, int d, float inv_tau,
                     float* __restrict__ e_out, float* __restrict__ row_sum, float* __restrict__ row_pos,
                     float* __restrict__ col_sum) {
  __shared__ float as[kKC][kTS + 4];
  __shared__ float bs[kKC][kTS + 4];
  __shared__ float colp[16][kTS];
  const int b = blockIdx.x;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const float* a_base = th + (int64_t)b * l * d;
  const float* b_base = oh + (int64_t)b * l * d;
  float acc[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
  for (int k0 = 0; k0 < d; k0 += kKC) {
#pragma unroll
    for (int i = 0; i < (kTS * kKC) / kThreads; ++i) {           // 8 elements per thread and operand
      const int idx = t + kThreads * i;
      const int row = idx / kKC, kk = idx - row * kKC;
      const bool ok = row < l && k0 + kk < d;
      as[kk][row] = ok ? __ldg(a_base + (int64_t)row * d + k0 + kk) : 0.f;
      bs[kk][row] = ok ? __ldg(b_base + (int64_t)row * d + k0 + kk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kKC; ++kk) {
      float av[8], bv[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) av[r] = as[kk][ty * 8 + r];
#pragma unroll
      for (int c = 0; c < 8; ++c) bv[c] = bs[kk][tx + 16 * c];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
    }
    __syncthreads();
  }
  const float shift = inv_tau;
  float cs[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) cs[c] = 0.f;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = ty * 8 + r;
    float rs = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = tx + 16 * c;
      const bool ok = i < l && j < l;
      const float sv = acc[r][c] * inv_tau;
      const float e = ok ? expf(sv - shift) : 0.f;
      if (ok) e_out[((int64_t)b * l + i) * l + j] = e;
      if (ok && i == j) row_pos[(int64_t)b * l + i] = sv;
      rs += e;
      cs[c] += e;
    }
    // the 16 threads of a row group are 16 consecutive lanes
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
    if (tx == 0 && i < l) row_sum[(int64_t)b * l + i] = rs;
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) colp[ty][tx + 16 * c] = cs[c];
  __syncthreads();
  if (t < l) {
    float sum = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) sum += colp[g][t];               // fixed order: deterministic
    col_sum[(int64_t)b * l + t] = sum;
  }
}

// d_th[b, i, c] = sum_j W_ij Ohat[j, c],  d_oh[b, j, c] = sum_i W_ij That[i, c],  W = E (a_i + b_j) - 2 [i == j]
// (identity targets: c_i = 1).  CTA = (64 feature columns, sample); W, the Ohat and That column slabs in shared memory.
constexpr int kDC = 64;
__global__ void __launch_bounds__(kThreads)
token_sim_bwd_kernel(const float* __restrict__ th, const float* __restrict__ oh, const float* __restrict__ e_in,
                     const float* __restrict__ a_row, const float* __restrict__ b_col, int l, int d,
                     float* __restrict__ d_th, float* __restrict__ d_oh) {
  extern __shared__ float sm[];
  float* w = sm;                                 // [kTS][kTS + 1]
  float* oc = w + kTS * (kTS + 1);               // [kTS][kDC]
  float* tc = oc + kTS * kDC;                    // [kTS][kDC]
  const int c0 = blockIdx.x * kDC, b = blockIdx.y;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int64_t base = (int64_t)b * l;
  for (int idx = t; idx < kTS * kTS; idx += kThreads) {
    const int i = idx / kTS, j = idx - i * kTS;
    float v = 0.f;
    if (i < l && j < l) v = e_in[(base + i) * l + j] * (__ldg(a_row + base + i) + __ldg(b_col + base + j)) - (i == j ? 2.f : 0.f);
    w[i * (kTS + 1) + j] = v;
  }
  for (int idx = t; idx < kTS * kDC; idx += kThreads) {
    const int j = idx / kDC, c = idx - j * kDC;
    const bool ok = j < l && c0 + c < d;
    oc[idx] = ok ? __ldg(oh + (base + j) * d + c0 + c) : 0.f;
    tc[idx] = ok ? __ldg(th + (base + j) * d + c0 + c) : 0.f;
  }
  __syncthreads();
  float acc[8][4];
  // rows of W
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  for (int j = 0; j < l; ++j) {
    const float4 o4 = *reinterpret_cast<const float4*>(oc + j * kDC + tx * 4);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float wv = w[(ty * 8 + r) * (kTS + 1) + j];
      acc[r][0] = fmaf(wv, o4.x, acc[r][0]); acc[r][1] = fmaf(wv, o4.y, acc[r][1]);
      acc[r][2] = fmaf(wv, o4.z, acc[r][2]); acc[r][3] = fmaf(wv, o4.w, acc[r][3]);
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = ty * 8 + r;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (i < l && c0 + tx * 4 + c < d) d_th[(base + i) * d + c0 + tx * 4 + c] = acc[r][c];
  }
  // columns of W
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  for (int i = 0; i < l; ++i) {
    const float4 t4 = *reinterpret_cast<const float4*>(tc + i * kDC + tx * 4);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float wv = w[i * (kTS + 1) + ty * 8 + r];
      acc[r][0] = fmaf(wv, t4.x, acc[r][0]); acc[r][1] = fmaf(wv, t4.y, acc[r][1]);
      acc[r][2] = fmaf(wv, t4.z, acc[r][2]); acc[r][3] = fmaf(wv, t4.w, acc[r][3]);
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int j = ty * 8 + r;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (j < l && c0 + tx * 4 + c < d) d_oh[(base + j) * d + c0 + tx * 4 + c] = acc[r][c];
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

extern "C" int evk_token_sim_fwd(const float* th, const float* oh, int64_t batch, int64_t l, int64_t d, float inv_tau,
                                 float* e_out, float* row_sum, float* row_pos, float* col_sum, evk_stream_t stream) {
  EVK_REQUIRE(th && oh && e_out && row_sum && row_pos && col_sum, "evk_token_sim_fwd: null pointer");
  EVK_REQUIRE(batch >= 1 && l >= 1 && l <= kTS && d >= 1, "evk_token_sim_fwd: needs 1 <= l <= %d tokens per sample", kTS);
  EVK_REQUIRE(inv_tau > 0.f && inv_tau <= EVK_MAX_INV_TAU, "evk_token_sim_fwd: 1/tau=%g outside (0, %g] (fixed-shift softmax)", inv_tau, EVK_MAX_INV_TAU);
  token_sim_fwd_kernel<<<(unsigned)batch, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(th, oh, (int)l, (int)d, inv_tau, e_out,
                                                                                         row_sum, row_pos, col_sum);
  EVK_CHECK_LAUNCH("token_sim_fwd");
  return EVK_OK;
}

extern "C" int evk_token_sim_bwd(const float* th, const float* oh, const float* e_in, const float* a_row, const float* b_col,
                                 int64_t batch, int64_t l, int64_t d, float* d_th, float* d_oh, evk_stream_t stream) {
  EVK_REQUIRE(th && oh && e_in && a_row && b_col && d_th && d_oh, "evk_token_sim_bwd: null pointer");
  EVK_REQUIRE(batch >= 1 && batch <= 65535 && l >= 1 && l <= kTS && d >= 1, "evk_token_sim_bwd: needs 1 <= l <= %d tokens per sample", kTS);
  const size_t smem = sizeof(float) * (size_t)(kTS * (kTS + 1) + 2 * kTS * kDC);
  EVK_CUDA(cudaFuncSetAttribute(token_sim_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  token_sim_bwd_kernel<<<dim3((unsigned)((d + kDC - 1) / kDC), (unsigned)batch), kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      th, oh, e_in, a_row, b_col, (int)l, (int)d, d_th, d_oh);
  EVK_CHECK_LAUNCH("token_sim_bwd");
  return EVK_OK;
}
