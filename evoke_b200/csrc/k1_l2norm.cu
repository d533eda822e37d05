// K1: fused row L2-normalise (+ bf16 hi/lo split, + row gather) and its backward.
// Replaces F.normalize(x, dim=-1, p=2) at models/model_pretrain_finetune_v0520.py:495-496, :436.
//
// HBM-bound: one warp per row, 2 x 128-bit loads per lane per step (a warp covers 1 KiB of a
// row per step, fully coalesced), row kept in registers between the sum-of-squares pass and
// the scale pass, 128-bit bf16 stores.  Algorithmic bytes per row (fp32 in, bf16 hi out):
// 4*d + 2*d (+2*d with the lo half, +4*d with the fp32 copy).
#include "evk_common.cuh"
#include "peer_sync.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// Fast path: fp32, unit column stride, 16-byte aligned rows, d % 8 == 0, d <= kIters*256.
template <int kIters>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l2norm_fwd_vec_kernel(const float* __restrict__ x, int64_t n_out, int d, int64_t stride_row,
                      const int32_t* __restrict__ gather, float* __restrict__ out_f32, int64_t ld_f32,
                      __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                      int64_t ld_bf16, float* __restrict__ norm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t r = warp0; r < n_out; r += nwarps) {
    const int64_t src = gather ? (int64_t)gather[r] : r;
    const float* xr = x + src * stride_row;
    float v[kIters][8];
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int c = (it * 32 + lane) * 8;
      if (c < d) {
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(xr + c));
        const float4 p1 = __ldg(reinterpret_cast<const float4*>(xr + c + 4));
        v[it][0] = p0.x; v[it][1] = p0.y; v[it][2] = p0.z; v[it][3] = p0.w;
        v[it][4] = p1.x; v[it][5] = p1.y; v[it][6] = p1.z; v[it][7] = p1.w;
#pragma unroll
        for (int e = 0; e < 8; ++e) ss = fmaf(v[it][e], v[it][e], ss);
      }
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    const float den = fmaxf(nrm, EVK_NORM_EPS);
    if (lane == 0) norm[r] = nrm;
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int c = (it * 32 + lane) * 8;
      if (c < d) {
        float h[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) h[e] = v[it][e] / den;
        if (out_f32) {
          float4* o = reinterpret_cast<float4*>(out_f32 + r * ld_f32 + c);
          o[0] = make_float4(h[0], h[1], h[2], h[3]);
          o[1] = make_float4(h[4], h[5], h[6], h[7]);
        }
        if (out_hi) {
          uint4 hi;
          hi.x = pack_bf16x2(h[0], h[1]); hi.y = pack_bf16x2(h[2], h[3]);
          hi.z = pack_bf16x2(h[4], h[5]); hi.w = pack_bf16x2(h[6], h[7]);
          *reinterpret_cast<uint4*>(out_hi + r * ld_bf16 + c) = hi;
          if (out_lo) {
            float l[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) l[e] = h[e] - __bfloat162float(__float2bfloat16_rn(h[e]));
            uint4 lo;
            lo.x = pack_bf16x2(l[0], l[1]); lo.y = pack_bf16x2(l[2], l[3]);
            lo.z = pack_bf16x2(l[4], l[5]); lo.w = pack_bf16x2(l[6], l[7]);
            *reinterpret_cast<uint4*>(out_lo + r * ld_bf16 + c) = lo;
          }
        }
      }
    }
  }
}

// Generic path: any dtype, any strides, any d (two passes over the row; the second hits L1/L2).
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l2norm_fwd_generic_kernel(const void* __restrict__ x, int x_dtype, int64_t n_out, int64_t d,
                          int64_t stride_row, int64_t stride_col, const int32_t* __restrict__ gather,
                          float* __restrict__ out_f32, int64_t ld_f32, __nv_bfloat16* __restrict__ out_hi,
                          __nv_bfloat16* __restrict__ out_lo, int64_t ld_bf16, float* __restrict__ norm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t r = warp0; r < n_out; r += nwarps) {
    const int64_t src = gather ? (int64_t)gather[r] : r;
    const int64_t base = src * stride_row;
    float ss = 0.f;
    for (int64_t c = lane; c < d; c += 32) {
      const float t = load_as_float(x, x_dtype, base + c * stride_col);
      ss = fmaf(t, t, ss);
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    const float den = fmaxf(nrm, EVK_NORM_EPS);
    if (lane == 0) norm[r] = nrm;
    for (int64_t c = lane; c < d; c += 32) {
      const float h = load_as_float(x, x_dtype, base + c * stride_col) / den;
      if (out_f32) out_f32[r * ld_f32 + c] = h;
      if (out_hi) {
        const __nv_bfloat16 hb = __float2bfloat16_rn(h);
        out_hi[r * ld_bf16 + c] = hb;
        if (out_lo) out_lo[r * ld_bf16 + c] = __float2bfloat16_rn(h - __bfloat162float(hb));
      }
    }
    if (out_hi) {                                  // pad columns [d, ld) are defined (zero)
      for (int64_t c = d + lane; c < ld_bf16; c += 32) {
        out_hi[r * ld_bf16 + c] = __float2bfloat16_rn(0.f);
        if (out_lo) out_lo[r * ld_bf16 + c] = __float2bfloat16_rn(0.f);
      }
    }
  }
}

// Backward: dx = scale * (g - xhat (xhat.g)) / ||x||   or   scale * g / eps when clamped.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l2norm_bwd_kernel(const void* __restrict__ x, int x_dtype, int64_t n_out, int64_t d, int64_t stride_row,
                  int64_t stride_col, const int32_t* __restrict__ gather, const float* __restrict__ norm,
                  const void* __restrict__ g, int g_dtype, int64_t ld_g, int n_parts, int64_t part_stride,
                  const float* __restrict__ scale_dev,
                  float scale_host, void* __restrict__ dx, int dx_dtype, int64_t ld_dx, int accumulate,
                  const int* __restrict__ error, const PeerSyncDev sync) {
  peer_sync_block(sync, blockIdx.x == 0);
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  float scale = scale_host * (scale_dev ? __ldg(scale_dev) : 1.f);
  if (error && *reinterpret_cast<const volatile int*>(error) != 0) scale = __int_as_float(0x7fc00000);   // poisoned step
  for (int64_t r = warp0; r < n_out; r += nwarps) {
    const int64_t src = gather ? (int64_t)gather[r] : r;
    const int64_t base = src * stride_row;
    const float nrm = norm[r];
    const bool clamped = nrm < EVK_NORM_EPS;
    const float den = fmaxf(nrm, EVK_NORM_EPS);
    auto gval = [&](int64_t c) {                   // upstream gradient = sum of the partial buffers (rank order)
      float v = load_as_float(g, g_dtype, r * ld_g + c);
      for (int p = 1; p < n_parts; ++p) v += load_as_float(g, g_dtype, p * part_stride + r * ld_g + c);
      return v;
    };
    float proj = 0.f;
    if (!clamped) {
      for (int64_t c = lane; c < d; c += 32) {
        const float h = load_as_float(x, x_dtype, base + c * stride_col) / den;
        proj = fmaf(h, gval(c), proj);
      }
      proj = warp_sum(proj);
    }
    for (int64_t c = lane; c < d; c += 32) {
      float o;
      const float gc = gval(c);
      if (clamped) {
        o = scale * (gc / EVK_NORM_EPS);
      } else {
        const float h = load_as_float(x, x_dtype, base + c * stride_col) / den;
        o = scale * ((gc - h * proj) / den);
      }
      const int64_t oi = src * ld_dx + c;
      if (accumulate && dx_dtype == EVK_DTYPE_F32) reinterpret_cast<float*>(dx)[oi] += o;
      else store_from_float(dx, dx_dtype, oi, o);
    }
  }
}

// Fast path: fp32 x / g / dx, unit column stride, 16-byte aligned rows, d % 4 == 0, d <= kIters*128.
// One pass: the row of x and g stays in registers between the projection and the update.
// Algorithmic bytes per row: 4d (x) + 4d (g) read, 4d (dx) written.
// kBf16G: the partial buffers hold bf16 (the compressed exchange of the sharded path): 4 values = one 64-bit load
// kParts: several partial buffers are summed (kept out of the single-buffer kernel: a runtime loop between the
// loads of the unrolled iterations serialises them, 27 -> 39 us at the bench shape)
template <int kIters, bool kBf16G, bool kParts>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l2norm_bwd_vec_kernel(const float* __restrict__ x, int64_t n_out, int d, int64_t stride_row,
                      const int32_t* __restrict__ gather, const float* __restrict__ norm,
                      const void* __restrict__ g, int64_t ld_g, int n_parts, int64_t part_stride,
                      const float* __restrict__ scale_dev,
                      float scale_host, float* __restrict__ dx, int64_t ld_dx, const int* __restrict__ error,
                      const PeerSyncDev sync) {
  peer_sync_block(sync, blockIdx.x == 0);         // sharded: the partial buffers were stored by the peers' contractions
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  float scale = scale_host * (scale_dev ? __ldg(scale_dev) : 1.f);
  // sharded path: a cross-GPU barrier of this step timed out -> NaN gradients instead of silently wrong ones
  if (error && *reinterpret_cast<const volatile int*>(error) != 0) scale = __int_as_float(0x7fc00000);
  for (int64_t r = warp0; r < n_out; r += nwarps) {
    const int64_t src = gather ? (int64_t)gather[r] : r;
    const float4* xr = reinterpret_cast<const float4*>(x + src * stride_row);
    auto gload = [&](int p, int c) -> float4 {       // 4 consecutive values of part p, row r
      if (kBf16G) {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(g) + p * part_stride + r * ld_g) + c);
        return make_float4(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u), __uint_as_float(t.y << 16),
                           __uint_as_float(t.y & 0xffff0000u));
      }
      return __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(g) + p * part_stride + r * ld_g) + c);
    };
    float4* dr = reinterpret_cast<float4*>(dx + src * ld_dx);
    const float nrm = norm[r];
    const bool clamped = nrm < EVK_NORM_EPS;
    const float den = fmaxf(nrm, EVK_NORM_EPS);
    const float inv_den = 1.f / den;
    float4 xv[kIters], gv[kIters];
    float proj = 0.f;
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int c = it * 32 + lane;
      if (c * 4 < d) {
        xv[it] = __ldg(xr + c);
        gv[it] = gload(0, c);
      }
    }
    if (kParts) {
      // partial dXhat buffers written by the ranks' K4b epilogues, summed in index order.  The loads of TWO parts
      // are in flight together (the round trips, not the bytes, bound this loop: 7 dependent rounds at 8 ranks)
      int p = 1;
      for (; p + 1 < n_parts; p += 2) {
        float4 t0[kIters], t1[kIters];
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
          const int c = it * 32 + lane;
          if (c * 4 < d) { t0[it] = gload(p, c); t1[it] = gload(p + 1, c); }
        }
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
          const int c = it * 32 + lane;
          if (c * 4 < d) {
            gv[it].x = (gv[it].x + t0[it].x) + t1[it].x; gv[it].y = (gv[it].y + t0[it].y) + t1[it].y;
            gv[it].z = (gv[it].z + t0[it].z) + t1[it].z; gv[it].w = (gv[it].w + t0[it].w) + t1[it].w;
          }
        }
      }
      for (; p < n_parts; ++p) {
        float4 t[kIters];
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
          const int c = it * 32 + lane;
          if (c * 4 < d) t[it] = gload(p, c);
        }
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
          const int c = it * 32 + lane;
          if (c * 4 < d) { gv[it].x += t[it].x; gv[it].y += t[it].y; gv[it].z += t[it].z; gv[it].w += t[it].w; }
        }
      }
    }
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int c = it * 32 + lane;
      if (c * 4 < d) {
        xv[it].x *= inv_den; xv[it].y *= inv_den; xv[it].z *= inv_den; xv[it].w *= inv_den;   // xhat (1 ulp of the forward's)
        proj = fmaf(xv[it].x, gv[it].x, proj);
        proj = fmaf(xv[it].y, gv[it].y, proj);
        proj = fmaf(xv[it].z, gv[it].z, proj);
        proj = fmaf(xv[it].w, gv[it].w, proj);
      }
    }
    proj = warp_sum(proj);
    const float s_over_den = clamped ? scale / EVK_NORM_EPS : scale / den;
    if (clamped) proj = 0.f;
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int c = it * 32 + lane;
      if (c * 4 < d) {
        float4 o;
        o.x = s_over_den * (gv[it].x - xv[it].x * proj);
        o.y = s_over_den * (gv[it].y - xv[it].y * proj);
        o.z = s_over_den * (gv[it].z - xv[it].z * proj);
        o.w = s_over_den * (gv[it].w - xv[it].w * proj);
        dr[c] = o;
      }
    }
  }
}

inline int grid_for_rows(int64_t n) {
  const int64_t blocks = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int64_t cap = (int64_t)evk_sm_count() * 8;  // 8 resident 256-thread CTAs per SM, one wave
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace

extern "C" int evk_l2norm_fwd(const void* x, int x_dtype, int64_t n_out, int64_t d, int64_t stride_row,
                              int64_t stride_col, const int32_t* gather, float* out_f32, int64_t ld_f32,
                              void* out_hi, void* out_lo, int64_t ld_bf16, float* norm,
                              evk_stream_t stream) {
  EVK_REQUIRE(x && norm, "evk_l2norm_fwd: x and norm must be non-null");
  EVK_REQUIRE(n_out >= 0 && d > 0, "evk_l2norm_fwd: bad shape n=%lld d=%lld", (long long)n_out, (long long)d);
  EVK_REQUIRE(x_dtype >= EVK_DTYPE_F32 && x_dtype <= EVK_DTYPE_F16, "evk_l2norm_fwd: bad dtype %d", x_dtype);
  EVK_REQUIRE(!out_lo || out_hi, "evk_l2norm_fwd: out_lo requires out_hi");
  EVK_REQUIRE(!out_hi || (ld_bf16 >= d && ld_bf16 % 8 == 0 && evk_aligned16(out_hi) && (!out_lo || evk_aligned16(out_lo))),
              "evk_l2norm_fwd: bf16 outputs need 16-byte aligned base and ld %% 8 == 0 (ld=%lld)", (long long)ld_bf16);
  EVK_REQUIRE(!out_f32 || ld_f32 >= d, "evk_l2norm_fwd: ld_f32 < d");
  if (n_out == 0) return EVK_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = grid_for_rows(n_out);
  const bool vec = x_dtype == EVK_DTYPE_F32 && stride_col == 1 && d % 8 == 0 && d <= 2048 &&
                   stride_row % 4 == 0 && evk_aligned16(x) &&
                   (!out_f32 || (ld_f32 % 4 == 0 && evk_aligned16(out_f32)));
  if (vec) {
    const float* xf = static_cast<const float*>(x);
    auto* hi = static_cast<__nv_bfloat16*>(out_hi);
    auto* lo = static_cast<__nv_bfloat16*>(out_lo);
    if (d <= 1024)
      l2norm_fwd_vec_kernel<4><<<grid, kWarpsPerBlock * 32, 0, s>>>(xf, n_out, (int)d, stride_row, gather, out_f32,
                                                                    ld_f32, hi, lo, ld_bf16, norm);
    else
      l2norm_fwd_vec_kernel<8><<<grid, kWarpsPerBlock * 32, 0, s>>>(xf, n_out, (int)d, stride_row, gather, out_f32,
                                                                    ld_f32, hi, lo, ld_bf16, norm);
  } else {
    l2norm_fwd_generic_kernel<<<grid, kWarpsPerBlock * 32, 0, s>>>(
        x, x_dtype, n_out, d, stride_row, stride_col, gather, out_f32, ld_f32,
        static_cast<__nv_bfloat16*>(out_hi), static_cast<__nv_bfloat16*>(out_lo), ld_bf16, norm);
  }
  EVK_CHECK_LAUNCH("l2norm_fwd");
  return EVK_OK;
}

extern "C" int evk_l2norm_bwd_parts(const void* x, int x_dtype, int64_t n_out, int64_t d, int64_t stride_row,
                                    int64_t stride_col, const int32_t* gather, const float* norm, const void* g,
                                    int g_dtype, int64_t ld_g, int n_parts, int64_t part_stride, const float* scale_dev,
                                    float scale_host, void* dx, int dx_dtype, int64_t ld_dx, int accumulate,
                                    const int* error, const evk_peer_sync_t* sync, evk_stream_t stream) {
  EVK_REQUIRE(x && norm && g && dx, "evk_l2norm_bwd: null pointer");
  PeerSyncDev ps;
  {
    int rc = peer_sync_from_host(sync, ps);
    if (rc != EVK_OK) return rc;
  }
  EVK_REQUIRE(n_parts >= 1 && (n_parts == 1 || part_stride >= n_out * ld_g), "evk_l2norm_bwd_parts: bad partial-buffer layout");
  EVK_REQUIRE(g_dtype == EVK_DTYPE_F32 || g_dtype == EVK_DTYPE_BF16, "evk_l2norm_bwd_parts: partial buffers must be fp32 or bf16");
  EVK_REQUIRE(n_out >= 0 && d > 0 && ld_g >= d && ld_dx >= d, "evk_l2norm_bwd: bad shape");
  EVK_REQUIRE(x_dtype >= EVK_DTYPE_F32 && x_dtype <= EVK_DTYPE_F16 && dx_dtype >= EVK_DTYPE_F32 &&
                  dx_dtype <= EVK_DTYPE_F16, "evk_l2norm_bwd: bad dtype");
  EVK_REQUIRE(!accumulate || dx_dtype == EVK_DTYPE_F32, "evk_l2norm_bwd: accumulate needs fp32 dx");
  EVK_REQUIRE(n_out > 0 || !sync, "evk_l2norm_bwd_parts: a folded sync needs at least one row");
  if (n_out == 0) return EVK_OK;
  const bool vec = x_dtype == EVK_DTYPE_F32 && dx_dtype == EVK_DTYPE_F32 && !accumulate && stride_col == 1 &&
                   d % 4 == 0 && d <= 2048 && stride_row % 4 == 0 && ld_g % 4 == 0 && ld_dx % 4 == 0 && part_stride % 4 == 0 &&
                   evk_aligned16(x) && evk_aligned16(g) && evk_aligned16(dx);
  const bool g16 = g_dtype == EVK_DTYPE_BF16;
  if (vec) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = grid_for_rows(n_out);
    const float* xf = static_cast<const float*>(x);
    float* df = static_cast<float*>(dx);
#define EVK_LAUNCH_BWD(IT, B16, PARTS)                                                                                    \
  l2norm_bwd_vec_kernel<IT, B16, PARTS><<<grid, kWarpsPerBlock * 32, 0, s>>>(xf, n_out, (int)d, stride_row, gather, norm, g, \
                                                                             ld_g, n_parts, part_stride, scale_dev,           \
                                                                             scale_host, df, ld_dx, error, ps)
    const bool multi = n_parts > 1;
    if (d <= 1024) {
      if (g16) { if (multi) EVK_LAUNCH_BWD(8, true, true); else EVK_LAUNCH_BWD(8, true, false); }
      else { if (multi) EVK_LAUNCH_BWD(8, false, true); else EVK_LAUNCH_BWD(8, false, false); }
    } else {
      if (g16) { if (multi) EVK_LAUNCH_BWD(16, true, true); else EVK_LAUNCH_BWD(16, true, false); }
      else { if (multi) EVK_LAUNCH_BWD(16, false, true); else EVK_LAUNCH_BWD(16, false, false); }
    }
#undef EVK_LAUNCH_BWD
    EVK_CHECK_LAUNCH("l2norm_bwd_vec");
    return EVK_OK;
  }
  l2norm_bwd_kernel<<<grid_for_rows(n_out), kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      x, x_dtype, n_out, d, stride_row, stride_col, gather, norm, g, g_dtype, ld_g, n_parts, part_stride, scale_dev, scale_host, dx,
      dx_dtype, ld_dx, accumulate, error, ps);
  EVK_CHECK_LAUNCH("l2norm_bwd");
  return EVK_OK;
}

extern "C" int evk_l2norm_bwd(const void* x, int x_dtype, int64_t n_out, int64_t d, int64_t stride_row,
                              int64_t stride_col, const int32_t* gather, const float* norm, const float* g,
                              int64_t ld_g, const float* scale_dev, float scale_host, void* dx, int dx_dtype,
                              int64_t ld_dx, int accumulate, evk_stream_t stream) {
  return evk_l2norm_bwd_parts(x, x_dtype, n_out, d, stride_row, stride_col, gather, norm, g, EVK_DTYPE_F32, ld_g, 1, 0, scale_dev,
                              scale_host, dx, dx_dtype, ld_dx, accumulate, nullptr, nullptr, stream);
}
