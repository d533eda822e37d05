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

#include <cooperative_groups.h>

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxP = 1024;                    // patches per sample held in shared memory

constexpr int kTile = 8;                       // tokens (or patches) per CTA: each patch / token row read is shared by 8

__global__ void __launch_bounds__(kThreads)
local_attend_fwd_kernel(const float* __restrict__ text, const float* __restrict__ image, int l_tokens, int p_tokens, int d,
                        float inv_sqrt_d, float* __restrict__ att, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
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
  extern __shared__ __align__(16) float sm[];
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
  extern __shared__ __align__(16) float sm[];
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

// ---- 128-bit variants (d % 4 == 0, 16-byte aligned rows): the kernels above are latency-bound on one or two scalar
// loads in flight per thread; here every thread has kU float4 loads in flight and the shared operands are read as
// float4.  Same CTA shape, same shared-memory layout, same order of the sums over patches / tokens.
constexpr int kU = 6;                          // float4 loads in flight per lane in the score loops (768 columns = 6 x 32 x 4)
constexpr int kUP = 7;                         // patches per batch of loads in the output loops (49 = 7 x 7)
constexpr int kUL = 4;                         // tokens per batch in the patch-gradient loop (2 loads each)

__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
}
__device__ __forceinline__ void axpy4(float s, const float4 v, float4& acc) {
  acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y); acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}

// scores of the CTA's 8 rows (rows_s: [kTile][d] in shared memory) against every patch row of the sample:
// sc[k][p] = scale * rows_s[k] . V[p]      (warp per patch, kU x 128-bit loads in flight per lane)
__device__ __forceinline__ void tile_scores_vec(const float4* __restrict__ rows_s4, const float4* __restrict__ v4, int p_tokens,
                                                int d4, float scale, float* __restrict__ sc, int warp, int lane) {
  for (int p = warp; p < p_tokens; p += kWarps) {
    float acc[kTile];
#pragma unroll
    for (int k = 0; k < kTile; ++k) acc[k] = 0.f;
    const float4* vr = v4 + (int64_t)p * d4;
    for (int c0 = 0; c0 < d4; c0 += 32 * kU) {
      float4 vv[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int c = c0 + u * 32 + lane;
        vv[u] = c < d4 ? __ldg(vr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int c = c0 + u * 32 + lane;
        if (c < d4) {
#pragma unroll
          for (int k = 0; k < kTile; ++k) acc[k] = dot4(rows_s4[k * d4 + c], vv[u], acc[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kTile; ++k) {
      const float s = warp_sum(acc[k]);
      if (lane == 0) sc[k * p_tokens + p] = s * scale;
    }
  }
}

// rows_out[l0 + k][c] (=|+=) sum_p w[k][p] V[p][c]   (thread per float4 column, kUP loads in flight, p ascending)
template <bool kAccumulate>
__device__ __forceinline__ void tile_weighted_rows_vec(const float* __restrict__ w, const float4* __restrict__ v4, int p_tokens,
                                                       int d4, float* __restrict__ rows_out, int64_t row0, int l0,
                                                       int l_tokens) {
  for (int c = threadIdx.x; c < d4; c += kThreads) {
    float4 acc[kTile];
#pragma unroll
    for (int k = 0; k < kTile; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p0 = 0; p0 < p_tokens; p0 += kUP) {
      float4 vv[kUP];
#pragma unroll
      for (int u = 0; u < kUP; ++u) vv[u] = p0 + u < p_tokens ? __ldg(v4 + (int64_t)(p0 + u) * d4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < kUP; ++u) {
        if (p0 + u < p_tokens) {
#pragma unroll
          for (int k = 0; k < kTile; ++k) axpy4(w[k * p_tokens + p0 + u], vv[u], acc[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kTile; ++k) {
      if (l0 + k < l_tokens) {
        float4* o = reinterpret_cast<float4*>(rows_out + (row0 + l0 + k) * (int64_t)(4 * d4)) + c;
        if (kAccumulate) {
          float4 t = *o;
          t.x += acc[k].x; t.y += acc[k].y; t.z += acc[k].z; t.w += acc[k].w;
          *o = t;
        } else {
          *o = acc[k];
        }
      }
    }
  }
}

__device__ __forceinline__ void load_tile_rows_vec(const float* __restrict__ src, int64_t row0, int l0, int l_tokens, int d4,
                                                   float4* __restrict__ dst4) {
  for (int i = threadIdx.x; i < kTile * d4; i += kThreads) {
    const int k = i / d4, c = i - k * d4;
    dst4[i] = (l0 + k < l_tokens) ? __ldg(reinterpret_cast<const float4*>(src + (row0 + l0 + k) * (int64_t)(4 * d4)) + c)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void __launch_bounds__(kThreads)
local_attend_fwd_vec_kernel(const float* __restrict__ text, const float* __restrict__ image, int l_tokens, int p_tokens, int d,
                            float inv_sqrt_d, float* __restrict__ att, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  float* t_s = sm;                              // [kTile][d]
  float* sc = sm + (size_t)kTile * d;           // [kTile][p_tokens]
  const int l0 = blockIdx.x * kTile, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d4 = d >> 2;
  const int64_t row0 = (int64_t)b * l_tokens;
  const float4* v4 = reinterpret_cast<const float4*>(image + (int64_t)b * p_tokens * d);
  load_tile_rows_vec(text, row0, l0, l_tokens, d4, reinterpret_cast<float4*>(t_s));
  __syncthreads();
  tile_scores_vec(reinterpret_cast<const float4*>(t_s), v4, p_tokens, d4, inv_sqrt_d, sc, warp, lane);
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
    float* a_row = att + (row0 + l0 + warp) * p_tokens;
    for (int p = lane; p < p_tokens; p += 32) {
      const float a = row[p] * inv_z;
      row[p] = a;
      if (ok) a_row[p] = a;
    }
  }
  __syncthreads();
  tile_weighted_rows_vec<false>(sc, v4, p_tokens, d4, out, row0, l0, l_tokens);
}

__global__ void __launch_bounds__(kThreads)
local_attend_bwd_token_vec_kernel(const float* __restrict__ image, const float* __restrict__ att,
                                  const float* __restrict__ d_out, int l_tokens, int p_tokens, int d, float inv_sqrt_d,
                                  float* __restrict__ ds, float* __restrict__ d_text) {
  extern __shared__ __align__(16) float sm[];
  float* g_s = sm;                              // [kTile][d]  dO of the 8 tokens
  float* da = sm + (size_t)kTile * d;           // [kTile][p_tokens]
  const int l0 = blockIdx.x * kTile, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d4 = d >> 2;
  const int64_t row0 = (int64_t)b * l_tokens;
  const float4* v4 = reinterpret_cast<const float4*>(image + (int64_t)b * p_tokens * d);
  load_tile_rows_vec(d_out, row0, l0, l_tokens, d4, reinterpret_cast<float4*>(g_s));
  __syncthreads();
  tile_scores_vec(reinterpret_cast<const float4*>(g_s), v4, p_tokens, d4, 1.f, da, warp, lane);
  __syncthreads();
  {                                                                // warp k <-> token k
    const bool ok = l0 + warp < l_tokens;
    const int64_t tok = row0 + l0 + warp;
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
  tile_weighted_rows_vec<true>(da, v4, p_tokens, d4, d_text, row0, l0, l_tokens);
}

__global__ void __launch_bounds__(kThreads)
local_attend_bwd_patch_vec_kernel(const float* __restrict__ text, const float* __restrict__ att, const float* __restrict__ ds,
                                  const float* __restrict__ d_out, int l_tokens, int p_tokens, int d,
                                  float* __restrict__ d_image) {
  extern __shared__ __align__(16) float sm[];
  float* a_s = sm;                              // [l_tokens][kTile] att[:, p0..p0+7]
  float* s_s = sm + (size_t)l_tokens * kTile;   // [l_tokens][kTile] ds[:, p0..p0+7]
  const int p0 = blockIdx.x * kTile, b = blockIdx.y;
  const int d4 = d >> 2;
  for (int i = threadIdx.x; i < l_tokens * kTile; i += kThreads) {
    const int l = i / kTile, k = i - l * kTile;
    const bool ok = p0 + k < p_tokens;
    const int64_t j = ((int64_t)b * l_tokens + l) * p_tokens + p0 + k;
    a_s[i] = ok ? att[j] : 0.f;
    s_s[i] = ok ? ds[j] : 0.f;
  }
  __syncthreads();
  const float4* t4 = reinterpret_cast<const float4*>(text + (int64_t)b * l_tokens * d);
  const float4* g4 = reinterpret_cast<const float4*>(d_out + (int64_t)b * l_tokens * d);
  for (int c = threadIdx.x; c < d4; c += kThreads) {
    float4 acc[kTile];
#pragma unroll
    for (int k = 0; k < kTile; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int lb = 0; lb < l_tokens; lb += kUL) {
      float4 gv[kUL], tv[kUL];
#pragma unroll
      for (int u = 0; u < kUL; ++u) {
        const bool ok = lb + u < l_tokens;
        gv[u] = ok ? __ldg(g4 + (int64_t)(lb + u) * d4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        tv[u] = ok ? __ldg(t4 + (int64_t)(lb + u) * d4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < kUL; ++u) {
        if (lb + u < l_tokens) {
          const int l = lb + u;
#pragma unroll
          for (int k = 0; k < kTile; ++k) {                       // acc = a * g + (s * t + acc), as in the scalar kernel
            axpy4(s_s[l * kTile + k], tv[u], acc[k]);
            axpy4(a_s[l * kTile + k], gv[u], acc[k]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kTile; ++k)
      if (p0 + k < p_tokens) reinterpret_cast<float4*>(d_image + ((int64_t)b * p_tokens + p0 + k) * d)[c] = acc[k];
  }
}

// ---- per-sample token-level InfoNCE (:518-525) for L <= 128: register-blocked fp32 kernels, one sample per CTA (row) ----
constexpr int kTS = 128;                       // tokens per sample held by one CTA
constexpr int kKC = 16;                        // feature chunk of the similarity product

// E[b, i, j] = exp(That_i . Ohat_j / tau - 1/tau), its row sums, column sums and the diagonal logits.
// 256 threads = 16 x 16; thread (ty, tx) owns rows 8 ty .. 8 ty + 7 and columns tx, tx + 16, .. (interleaved: the
// shared-memory reads of the column operand and the stores of E are then conflict-free / coalesced).
template <bool kVec>
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
  // kVec (d % 4 == 0, 16-byte aligned rows): the next feature chunk is fetched into registers (2 x 128 bits per
  // operand and thread) while the current one is multiplied, so the global-load latency is hidden behind the
  // 1024 FMAs per thread of a chunk instead of being exposed 48 times; same order of accumulation.
  float4 pa[2], pb[2];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = t + kThreads * i;                          // float4 index in the [kTS rows][4 float4] chunk
      const int row = idx >> 2, q4 = idx & 3;
      const bool ok = row < l && k0 + 4 * q4 < d;
      pa[i] = ok ? __ldg(reinterpret_cast<const float4*>(a_base + (int64_t)row * d + k0) + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
      pb[i] = ok ? __ldg(reinterpret_cast<const float4*>(b_base + (int64_t)row * d + k0) + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if (kVec) fetch(0);
  for (int k0 = 0; k0 < d; k0 += kKC) {
    if (kVec) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int idx = t + kThreads * i;
        const int row = idx >> 2, kk = (idx & 3) * 4;
        as[kk][row] = pa[i].x; as[kk + 1][row] = pa[i].y; as[kk + 2][row] = pa[i].z; as[kk + 3][row] = pa[i].w;
        bs[kk][row] = pb[i].x; bs[kk + 1][row] = pb[i].y; bs[kk + 2][row] = pb[i].z; bs[kk + 3][row] = pb[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < (kTS * kKC) / kThreads; ++i) {         // 8 elements per thread and operand
        const int idx = t + kThreads * i;
        const int row = idx / kKC, kk = idx - row * kKC;
        const bool ok = row < l && k0 + kk < d;
        as[kk][row] = ok ? __ldg(a_base + (int64_t)row * d + k0 + kk) : 0.f;
        bs[kk][row] = ok ? __ldg(b_base + (int64_t)row * d + k0 + kk) : 0.f;
      }
    }
    __syncthreads();
    if (kVec && k0 + kKC < d) fetch(k0 + kKC);
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

// The same for a CLUSTER of kTSC CTAs per sample (d % 4 == 0): one CTA per sample leaves 116 of 148 SMs idle at the
// reference's batch (B = 32).  CTA r multiplies the r-th quarter of the feature chunks into a partial L x L tile, the
// partial tiles are summed through distributed shared memory in rank order (deterministic), and CTA r finishes rows
// [32 r, 32 r + 32) - exp, row sums, diagonal, column partials; rank 0 adds the column partials of the four CTAs.
constexpr int kTSC = 4;
namespace cg = cooperative_groups;

__global__ void __cluster_dims__(kTSC, 1, 1) __launch_bounds__(kThreads)
token_sim_fwd_cluster_kernel(const float* __restrict__ th, const float* __restrict__ oh, int l, int d, float inv_tau,
                             float* __restrict__ e_out, float* __restrict__ row_sum, float* __restrict__ row_pos,
                             float* __restrict__ col_sum) {
  __shared__ float as[kKC][kTS + 4];
  __shared__ float bs[kKC][kTS + 4];
  __shared__ float colp[4][kTS];
  __shared__ float colq[kTS];
  extern __shared__ __align__(16) float sm[];                    // part[kTS][kTS]: this CTA's partial tile
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / kTSC;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const float* a_base = th + (int64_t)b * l * d;
  const float* b_base = oh + (int64_t)b * l * d;
  float acc[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
  const int nchunk = (d + kKC - 1) / kKC;
  const int k_lo = (nchunk * rank / kTSC) * kKC, k_hi = (nchunk * (rank + 1) / kTSC) * kKC;
  float4 pa[2], pb[2];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = t + kThreads * i;
      const int row = idx >> 2, q4 = idx & 3;
      const bool ok = row < l && k0 + 4 * q4 < d;
      pa[i] = ok ? __ldg(reinterpret_cast<const float4*>(a_base + (int64_t)row * d + k0) + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
      pb[i] = ok ? __ldg(reinterpret_cast<const float4*>(b_base + (int64_t)row * d + k0) + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if (k_lo < k_hi) fetch(k_lo);
  for (int k0 = k_lo; k0 < k_hi; k0 += kKC) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = t + kThreads * i;
      const int row = idx >> 2, kk = (idx & 3) * 4;
      as[kk][row] = pa[i].x; as[kk + 1][row] = pa[i].y; as[kk + 2][row] = pa[i].z; as[kk + 3][row] = pa[i].w;
      bs[kk][row] = pb[i].x; bs[kk + 1][row] = pb[i].y; bs[kk + 2][row] = pb[i].z; bs[kk + 3][row] = pb[i].w;
    }
    __syncthreads();
    if (k0 + kKC < k_hi) fetch(k0 + kKC);
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
  float* part = sm;
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) part[(ty * 8 + r) * kTS + tx + 16 * c] = acc[r][c];
  cluster.sync();                                                  // every CTA's partial tile is complete
  // rows [32 rank, 32 rank + 32) = ty in [4 rank, 4 rank + 4) = warps 2 rank and 2 rank + 1 (whole warps)
  if ((ty >> 2) == rank) {
    const float* ps[kTSC];
#pragma unroll
    for (int sr = 0; sr < kTSC; ++sr) ps[sr] = cluster.map_shared_rank(part, sr);
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
        const int o = i * kTS + j;
        const float sacc = ((ps[0][o] + ps[1][o]) + ps[2][o]) + ps[3][o];   // fixed order: deterministic
        const bool ok = i < l && j < l;
        const float sv = sacc * inv_tau;
        const float e = ok ? expf(sv - shift) : 0.f;
        if (ok) e_out[((int64_t)b * l + i) * l + j] = e;
        if (ok && i == j) row_pos[(int64_t)b * l + i] = sv;
        rs += e;
        cs[c] += e;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
      if (tx == 0 && i < l) row_sum[(int64_t)b * l + i] = rs;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) colp[ty & 3][tx + 16 * c] = cs[c];
  }
  __syncthreads();
  if (t < kTS) colq[t] = ((colp[0][t] + colp[1][t]) + colp[2][t]) + colp[3][t];
  cluster.sync();                                                  // column partials of the four row quarters
  if (rank == 0 && t < l) {
    float sum = 0.f;
#pragma unroll
    for (int sr = 0; sr < kTSC; ++sr) sum += cluster.map_shared_rank(colq, sr)[t];
    col_sum[(int64_t)b * l + t] = sum;
  }
  cluster.sync();                                                  // a CTA's shared memory must outlive the peers' reads
}

// d_th[b, i, c] = sum_j W_ij Ohat[j, c],  d_oh[b, j, c] = sum_i W_ij That[i, c],  W = E (a_i + b_j) - 2 [i == j]
// (identity targets: c_i = 1).  CTA = (64 feature columns, sample); W, the Ohat and That column slabs in shared memory.
constexpr int kDC = 64;
__global__ void __launch_bounds__(kThreads)
token_sim_bwd_kernel(const float* __restrict__ th, const float* __restrict__ oh, const float* __restrict__ e_in,
                     const float* __restrict__ a_row, const float* __restrict__ b_col, int l, int d,
                     float* __restrict__ d_th, float* __restrict__ d_oh) {
  extern __shared__ __align__(16) float sm[];
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
  const bool vec = d % 4 == 0 && evk_aligned16(text) && evk_aligned16(image) && evk_aligned16(out);
  auto kern = vec ? local_attend_fwd_vec_kernel : local_attend_fwd_kernel;
  EVK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3((unsigned)((l + kTile - 1) / kTile), (unsigned)batch), kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
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
  const bool vec = d % 4 == 0 && evk_aligned16(text) && evk_aligned16(image) && evk_aligned16(d_out) && evk_aligned16(d_text) &&
                   evk_aligned16(d_image);
  auto k_tok = vec ? local_attend_bwd_token_vec_kernel : local_attend_bwd_token_kernel;
  EVK_CUDA(cudaFuncSetAttribute(k_tok, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  k_tok<<<dim3((unsigned)((l + kTile - 1) / kTile), (unsigned)batch), kThreads, smem1, s>>>(
      image, att, d_out, (int)l, (int)p, (int)d, 1.f / sqrtf((float)d), ds, d_text);
  EVK_CHECK_LAUNCH("local_attend_bwd_token");
  const size_t smem2 = sizeof(float) * (size_t)(2 * l * kTile);
  auto k_pat = vec ? local_attend_bwd_patch_vec_kernel : local_attend_bwd_patch_kernel;
  EVK_CUDA(cudaFuncSetAttribute(k_pat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  k_pat<<<dim3((unsigned)((p + kTile - 1) / kTile), (unsigned)batch), kThreads, smem2, s>>>(text, att, ds, d_out, (int)l, (int)p,
                                                                                          (int)d, d_image);
  EVK_CHECK_LAUNCH("local_attend_bwd_patch");
  return EVK_OK;
}

extern "C" int evk_token_sim_fwd(const float* th, const float* oh, int64_t batch, int64_t l, int64_t d, float inv_tau,
                                 float* e_out, float* row_sum, float* row_pos, float* col_sum, evk_stream_t stream) {
  EVK_REQUIRE(th && oh && e_out && row_sum && row_pos && col_sum, "evk_token_sim_fwd: null pointer");
  EVK_REQUIRE(batch >= 1 && l >= 1 && l <= kTS && d >= 1, "evk_token_sim_fwd: needs 1 <= l <= %d tokens per sample", kTS);
  EVK_REQUIRE(inv_tau > 0.f && inv_tau <= EVK_MAX_INV_TAU, "evk_token_sim_fwd: 1/tau=%g outside (0, %g] (fixed-shift softmax)", inv_tau, EVK_MAX_INV_TAU);
  const bool vec = d % 4 == 0 && evk_aligned16(th) && evk_aligned16(oh);
  static const bool use_cluster = [] { const char* e = getenv("EVK_F1_CLUSTER"); return !(e && e[0] == '0'); }();
  if (vec && use_cluster) {                                       // kTSC CTAs per sample, partial tiles summed through DSMEM
    const size_t smem = sizeof(float) * (size_t)kTS * kTS;
    EVK_CUDA(cudaFuncSetAttribute(token_sim_fwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    token_sim_fwd_cluster_kernel<<<(unsigned)(batch * kTSC), kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        th, oh, (int)l, (int)d, inv_tau, e_out, row_sum, row_pos, col_sum);
    EVK_CHECK_LAUNCH("token_sim_fwd_cluster");
    return EVK_OK;
  }
  auto kern = vec ? token_sim_fwd_kernel<true> : token_sim_fwd_kernel<false>;
  kern<<<(unsigned)batch, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(th, oh, (int)l, (int)d, inv_tau, e_out, row_sum, row_pos,
                                                                         col_sum);
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
