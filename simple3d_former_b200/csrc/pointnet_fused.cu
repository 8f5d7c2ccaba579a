// Set-abstraction / feature-propagation layers of the point tokenizer (reference data/pointnet_util.py:99-138, 220-244,
// 381-420; models/3DViT/model.py:33-72) as HBM-bound sm_100a kernels around the tcgen05 GEMM.
//
// The reference materialises the grouped tensor [B, S, K, 3 + Cf], runs Conv2d(1x1) -> BatchNorm2d -> ReLU twice on it
// (NCHW, fp32) and takes the max over the K neighbours. Here
//   * layer 1 is linear in its input, so it is evaluated per POINT (u = f W1f^T + b1, a [B*N, C1] GEMM, K x fewer rows
//     than the grouped tensor) and the per-neighbour value is z1 = u[idx] + W1x (xyz[idx] - centre), with the centred
//     coordinates handled in fp32 on the CUDA cores (exact differences, no bf16 cancellation);
//   * BatchNorm (training: batch statistics) needs one statistics pass per layer; the passes re-gather u (L2-resident)
//     instead of storing z1;
//   * layer 2 is one tensor-core GEMM over the bf16 activations, followed by one pass that produces the per-group
//     max / min (BatchNorm's scale may be negative, so the pooled value is max or min of z2) and the statistics;
//   * backward mirrors this: BatchNorm backward needs sum(dy), sum(dy * zhat) before dz, one statistics pass each.
// Statistics are reduced through per-CTA partial sums and a fixed-order fp64 finalize: forward results are bitwise
// reproducible run to run. Only scatter-adds of gradients use atomics.
//
// Work split of the row kernels: a warp owns one group (or row) and a chunk of 128 channels: lane l holds the channel
// quad 4l..4l+3 of the chunk, so every load / store / reduction is a 16-byte (fp32) or 8-byte (bf16) vector access
// and a warp touches 512 (256) contiguous bytes. All warps of a CTA work on the same chunk, so per-channel accumulators reduce inside the CTA.
#include "kernels.h"

namespace s3d {

namespace {

constexpr int kChunk = 128;   // channels per warp pass
constexpr int kWarps = 8;     // warps per CTA

struct SaGroup {
  const float* uf;        // [B*N, C1] per-point linear part (bias included)
  const float* xyz;       // [B, N, 3]
  const float* cxyz;      // [B, S, 3] group centres
  const long long* idx;   // [B, S, K]
  const float* w1;        // [C1, ldw]; columns 0..2 multiply the centred coordinates
  int ldw;
  int B, N, S, K, C1;
};

struct SaExtra {
  const float* scale;     // gamma * rstd           (modes 1, 3)
  const float* shift;     // beta - mean * scale    (mode 1)
  const float* mean;      // modes 2, 3
  const float* rstd;      // modes 2, 3
  const float* m1;        // mode 3: mean(dy)
  const float* m2;        // mode 3: mean(dy * zhat)
  __nv_bfloat16* a1_out;  // mode 1: [R, C1]
  const __nv_bfloat16* a1;    // modes 2, 3
  const __nv_bfloat16* da1;   // modes 2, 3
  float* duf;             // mode 3: [B*N, C1], zeroed by the caller
  float* partials;        // modes 0, 2: [P, 2, C1]; mode 3: [P, 3, C1]
};

__device__ __forceinline__ float4 ld4(const float* p, bool ok) {
  return ok ? *reinterpret_cast<const float4*>(p) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void ld4_bf16(const __nv_bfloat16* p, bool ok, float (&v)[4]) {
  if (ok) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else {
    v[0] = v[1] = v[2] = v[3] = 0.f;
  }
}
__device__ __forceinline__ void st4_bf16(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ void ldc4(const float* p, int c, bool ok, float (&v)[4]) {
  const float4 t = ld4(p + c, ok);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}

// CTA reduction of per-lane channel accumulators acc[set][4] (lane l owns channels 4l..4l+3 of the chunk) and write of
// one partial slot.
template <int NSETS>
__device__ __forceinline__ void cta_reduce_write(float (&acc)[NSETS][4], float* red /*[kWarps][NSETS][kChunk]*/,
                                                 float* partial_slot /*[NSETS][C]*/, int C, int chunk_base) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < NSETS; ++s) {
    float* r = red + ((size_t)wib * NSETS + s) * kChunk;
    *reinterpret_cast<float4*>(r + 4 * lane) = make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NSETS * kChunk; i += blockDim.x) {
    const int s = i / kChunk, c = i % kChunk;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += red[((size_t)w * NSETS + s) * kChunk + c];
    if (chunk_base + c < C) partial_slot[(size_t)s * C + chunk_base + c] = t;
  }
}

// MODE 0: sum z1, sum z1^2                     (BatchNorm-1 forward statistics)
// MODE 1: A1 = bf16(relu(scale * z1 + shift))   (layer-2 GEMM operand)
// MODE 2: sum dy1, sum dy1 * zhat1              (BatchNorm-1 backward statistics; dy1 = dA1 where A1 > 0)
// MODE 3: dz1 = scale * (dy1 - m1 - zhat1 * m2); duf[idx] += dz1; sum dz1 * (centred xyz)  (dW1x)
template <int MODE>
__global__ void __launch_bounds__(kWarps * 32, MODE == 3 ? 3 : 4) sa_group_kernel(SaGroup a, SaExtra e) {
  pdl_prologue();
  constexpr int NSETS = (MODE == 3) ? 3 : 2;
  __shared__ __align__(16) float red[(MODE == 1) ? 4 : kWarps * NSETS * kChunk];
  const int nchunks = (a.C1 + kChunk - 1) / kChunk;
  const int q = blockIdx.x % nchunks;
  const int part = blockIdx.x / nchunks;
  const int nparts = gridDim.x / nchunks;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int c0 = q * kChunk + 4 * lane;
  const bool ok = c0 < a.C1;

  float wx[4][3];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int t = 0; t < 3; ++t) wx[j][t] = ok ? a.w1[(size_t)(c0 + j) * a.ldw + t] : 0.f;
  float sc[4] = {0, 0, 0, 0}, sh[4] = {0, 0, 0, 0}, mu[4] = {0, 0, 0, 0}, rs[4] = {0, 0, 0, 0};
  float g1[4] = {0, 0, 0, 0}, g2[4] = {0, 0, 0, 0};
  if (MODE == 1 || MODE == 3) ldc4(e.scale, c0, ok, sc);
  if (MODE == 1) ldc4(e.shift, c0, ok, sh);
  if (MODE >= 2) { ldc4(e.mean, c0, ok, mu); ldc4(e.rstd, c0, ok, rs); }
  if (MODE == 3) { ldc4(e.m1, c0, ok, g1); ldc4(e.m2, c0, ok, g2); }
  float acc[NSETS][4];
#pragma unroll
  for (int s = 0; s < NSETS; ++s)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[s][j] = 0.f;

  const long long G = (long long)a.B * a.S;
  for (long long g = (long long)part * kWarps + wib; g < G; g += (long long)nparts * kWarps) {
    const long long b = g / a.S;
    long long my_i = 0;
    float dx = 0.f, dy = 0.f, dz = 0.f;
    if (lane < a.K) {
      my_i = a.idx[g * a.K + lane];
      const float* p = a.xyz + ((size_t)b * a.N + (size_t)my_i) * 3;
      const float* c = a.cxyz + (size_t)g * 3;
      dx = p[0] - c[0];
      dy = p[1] - c[1];
      dz = p[2] - c[2];
    }
#pragma unroll 4
    for (int k = 0; k < a.K; ++k) {
      const long long ik = __shfl_sync(0xffffffffu, my_i, k);
      const float x = __shfl_sync(0xffffffffu, dx, k);
      const float y = __shfl_sync(0xffffffffu, dy, k);
      const float w = __shfl_sync(0xffffffffu, dz, k);
      const size_t prow = (size_t)b * a.N + (size_t)ik;
      const float4 u = ld4(a.uf + prow * a.C1 + c0, ok);
      float z[4];
      z[0] = u.x + (wx[0][0] * x + wx[0][1] * y + wx[0][2] * w);
      z[1] = u.y + (wx[1][0] * x + wx[1][1] * y + wx[1][2] * w);
      z[2] = u.z + (wx[2][0] * x + wx[2][1] * y + wx[2][2] * w);
      z[3] = u.w + (wx[3][0] * x + wx[3][1] * y + wx[3][2] * w);
      const size_t row = (size_t)g * a.K + k;
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[0][j] += z[j];
          acc[1][j] += z[j] * z[j];
        }
      } else if (MODE == 1) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = fmaxf(fmaf(sc[j], z[j], sh[j]), 0.f);
        if (ok) st4_bf16(e.a1_out + row * a.C1 + c0, v);
      } else {
        float av[4], d[4], zh[4];
        ld4_bf16(e.a1 + row * a.C1 + c0, ok, av);
        ld4_bf16(e.da1 + row * a.C1 + c0, ok, d);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          d[j] = av[j] > 0.f ? d[j] : 0.f;
          zh[j] = (z[j] - mu[j]) * rs[j];
        }
        if (MODE == 2) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[0][j] += d[j];
            acc[1][j] += d[j] * zh[j];
          }
        } else {
          float t[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            t[j] = sc[j] * (d[j] - g1[j] - zh[j] * g2[j]);
            acc[0][j] += t[j] * x;
            acc[1][j] += t[j] * y;
            acc[2][j] += t[j] * w;
          }
          if (ok) red_add_f32x4(e.duf + prow * a.C1 + c0, t[0], t[1], t[2], t[3]);
        }
      }
    }
  }
  if (MODE != 1) cta_reduce_write<NSETS>(acc, red, e.partials + (size_t)part * NSETS * a.C1, a.C1, q * kChunk);
}

// z2 [G*K, C] fp32 -> per group and channel: max / min over the K rows and the row (k) where they occur (first on
// ties), plus the BatchNorm statistics sum z2, sum z2^2 over ALL rows.
__global__ void __launch_bounds__(kWarps * 32, 4) sa_group_reduce_kernel(const float* __restrict__ z2, long long G, int K,
                                                                     int C, float* __restrict__ zmax,
                                                                     float* __restrict__ zmin,
                                                                     unsigned char* __restrict__ kmax,
                                                                     unsigned char* __restrict__ kmin,
                                                                     float* __restrict__ partials) {
  pdl_prologue();
  __shared__ __align__(16) float red[kWarps * 2 * kChunk];
  const int nchunks = (C + kChunk - 1) / kChunk;
  const int q = blockIdx.x % nchunks, part = blockIdx.x / nchunks, nparts = gridDim.x / nchunks;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int c0 = q * kChunk + 4 * lane;
  const bool ok = c0 < C;
  float acc[2][4];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[s][j] = 0.f;
  for (long long g = (long long)part * kWarps + wib; g < G; g += (long long)nparts * kWarps) {
    float mx[4], mn[4];
    int ix[4], in_[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { mx[j] = -INFINITY; mn[j] = INFINITY; ix[j] = 0; in_[j] = 0; }
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      float v[4];
      ldc4(z2 + ((size_t)g * K + k) * C, c0, ok, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[0][j] += v[j];
        acc[1][j] += v[j] * v[j];
        if (v[j] > mx[j]) { mx[j] = v[j]; ix[j] = k; }
        if (v[j] < mn[j]) { mn[j] = v[j]; in_[j] = k; }
      }
    }
    if (ok) {
      const size_t o = (size_t)g * C + c0;
      *reinterpret_cast<float4*>(zmax + o) = make_float4(mx[0], mx[1], mx[2], mx[3]);
      *reinterpret_cast<float4*>(zmin + o) = make_float4(mn[0], mn[1], mn[2], mn[3]);
      *reinterpret_cast<uchar4*>(kmax + o) = make_uchar4((unsigned char)ix[0], (unsigned char)ix[1],
                                                         (unsigned char)ix[2], (unsigned char)ix[3]);
      *reinterpret_cast<uchar4*>(kmin + o) = make_uchar4((unsigned char)in_[0], (unsigned char)in_[1],
                                                         (unsigned char)in_[2], (unsigned char)in_[3]);
    }
  }
  cta_reduce_write<2>(acc, red, partials + (size_t)part * 2 * C, C, q * kChunk);
}

// out = relu(scale * zsel + shift) with zsel = scale >= 0 ? zmax : zmin  (max_k relu(bn(z_k)) == relu(bn(max or min)))
__global__ void sa_pool_select_kernel(const float* __restrict__ zmax, const float* __restrict__ zmin,
                                      const unsigned char* __restrict__ kmax, const unsigned char* __restrict__ kmin,
                                      const float* __restrict__ scale, const float* __restrict__ shift,
                                      float* __restrict__ out, float* __restrict__ zsel,
                                      unsigned char* __restrict__ ksel, long long n, int C) {
  pdl_prologue();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float s = scale[c];
    const bool up = s >= 0.f;
    const float z = up ? zmax[i] : zmin[i];
    zsel[i] = z;
    ksel[i] = up ? kmax[i] : kmin[i];
    out[i] = fmaxf(fmaf(s, z, shift[c]), 0.f);
  }
}

// dz2[r, c] = scale * (dy - m1 - zhat * m2), dy = dout[g, c] at the pooled row (k == ksel) when its output was > 0.
__global__ void __launch_bounds__(kWarps * 32, 4) sa_dz2_expand_kernel(
    const float* __restrict__ z2, const float* __restrict__ dout, const float* __restrict__ zsel,
    const unsigned char* __restrict__ ksel, const float* __restrict__ scale, const float* __restrict__ shift,
    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ m1,
    const float* __restrict__ m2, __nv_bfloat16* __restrict__ dz2, long long G, int K, int C) {
  pdl_prologue();
  const int nchunks = (C + kChunk - 1) / kChunk;
  const int q = blockIdx.x % nchunks, part = blockIdx.x / nchunks, nparts = gridDim.x / nchunks;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int c0 = q * kChunk + 4 * lane;
  const bool ok = c0 < C;
  float sc[4], sh[4], mu[4], rs[4], g1[4], g2[4];
  ldc4(scale, c0, ok, sc);
  ldc4(shift, c0, ok, sh);
  ldc4(mean, c0, ok, mu);
  ldc4(rstd, c0, ok, rs);
  ldc4(m1, c0, ok, g1);
  ldc4(m2, c0, ok, g2);
  for (long long g = (long long)part * kWarps + wib; g < G; g += (long long)nparts * kWarps) {
    const size_t o = (size_t)g * C + c0;
    float dsel[4], zs[4], dy[4];
    ldc4(dout + (size_t)g * C, c0, ok, dsel);
    ldc4(zsel + (size_t)g * C, c0, ok, zs);
    int ks[4] = {-1, -1, -1, -1};
    if (ok) {
      const uchar4 u = *reinterpret_cast<const uchar4*>(ksel + o);
      ks[0] = u.x; ks[1] = u.y; ks[2] = u.z; ks[3] = u.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) dy[j] = (fmaf(sc[j], zs[j], sh[j]) > 0.f) ? dsel[j] : 0.f;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const size_t row = (size_t)g * K + k;
      float v[4], t[4];
      ldc4(z2 + row * C, c0, ok, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float zh = (v[j] - mu[j]) * rs[j];
        t[j] = sc[j] * ((k == ks[j] ? dy[j] : 0.f) - g1[j] - zh * g2[j]);
      }
      if (ok) st4_bf16(dz2 + row * C + c0, t);
    }
  }
}

// Row kernels for BatchNorm over a plain [R, C] fp32 matrix (Linear -> BatchNorm1d -> ReLU of TransitionUp).
// MODE 0: sum z, sum z^2.  MODE 1: sum dy, sum dy * zhat with dy = dout where scale * z + shift > 0.
template <int MODE>
__global__ void __launch_bounds__(kWarps * 32, 4) bn_rows_stats_kernel(const float* __restrict__ z,
                                                                   const float* __restrict__ dout,
                                                                   const float* __restrict__ scale,
                                                                   const float* __restrict__ shift,
                                                                   const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd, long long R, int C,
                                                                   float* __restrict__ partials) {
  pdl_prologue();
  __shared__ __align__(16) float red[kWarps * 2 * kChunk];
  const int nchunks = (C + kChunk - 1) / kChunk;
  const int q = blockIdx.x % nchunks, part = blockIdx.x / nchunks, nparts = gridDim.x / nchunks;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int c0 = q * kChunk + 4 * lane;
  const bool ok = c0 < C;
  float sc[4] = {0, 0, 0, 0}, sh[4] = {0, 0, 0, 0}, mu[4] = {0, 0, 0, 0}, rs[4] = {0, 0, 0, 0};
  if (MODE == 1) {
    ldc4(scale, c0, ok, sc);
    ldc4(shift, c0, ok, sh);
    ldc4(mean, c0, ok, mu);
    ldc4(rstd, c0, ok, rs);
  }
  float acc[2][4];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[s][j] = 0.f;
#pragma unroll 4
  for (long long r = (long long)part * kWarps + wib; r < R; r += (long long)nparts * kWarps) {
    float v[4];
    ldc4(z + (size_t)r * C, c0, ok, v);
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[0][j] += v[j]; acc[1][j] += v[j] * v[j]; }
    } else {
      float d[4];
      ldc4(dout + (size_t)r * C, c0, ok, d);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float dy = (fmaf(sc[j], v[j], sh[j]) > 0.f) ? d[j] : 0.f;
        acc[0][j] += dy;
        acc[1][j] += dy * ((v[j] - mu[j]) * rs[j]);
      }
    }
  }
  cta_reduce_write<2>(acc, red, partials + (size_t)part * 2 * C, C, q * kChunk);
}

// Forward finalize: batch mean / biased variance from the partial sums (fixed order, fp64), BatchNorm affine as
// scale = gamma * rstd, shift = beta - mean * scale, running statistics update with the unbiased variance.
// One warp per channel: lane l sums partial slots l, l+32, ... in fp64, then a fixed shuffle tree (deterministic).
__device__ __forceinline__ void finalize_sums(const float* __restrict__ partials, int P, int C, int c, double& s,
                                              double& ss) {
  const int lane = threadIdx.x & 31;
  s = 0.0;
  ss = 0.0;
  for (int p = lane; p < P; p += 32) {
    s += (double)partials[((size_t)p * 2) * C + c];
    ss += (double)partials[((size_t)p * 2 + 1) * C + c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
}

__global__ void bn_finalize_fwd_kernel(const float* __restrict__ partials, int P, int C, double count,
                                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                       float momentum, float* __restrict__ running_mean,
                                       float* __restrict__ running_var, float* __restrict__ mean,
                                       float* __restrict__ rstd, float* __restrict__ scale, float* __restrict__ shift) {
  pdl_prologue();
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= C) return;  // warp-uniform
  double s, ss;
  finalize_sums(partials, P, C, c, s, ss);
  if ((threadIdx.x & 31) != 0) return;
  const double m = s / count;
  double var = ss / count - m * m;
  if (var < 0.0) var = 0.0;
  const float r = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma != nullptr ? gamma[c] : 1.f;
  const float b = beta != nullptr ? beta[c] : 0.f;
  mean[c] = (float)m;
  rstd[c] = r;
  scale[c] = g * r;
  shift[c] = b - (float)m * (g * r);
  if (running_mean != nullptr) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
  if (running_var != nullptr) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// Backward finalize: dgamma = sum dy * zhat, dbeta = sum dy; m1 = dbeta / count, m2 = dgamma / count in training mode
// (batch statistics take part in the gradient), zero in eval mode (running statistics are constants).
__global__ void bn_finalize_bwd_kernel(const float* __restrict__ partials, int P, int C, double count, int training,
                                       float* __restrict__ m1, float* __restrict__ m2, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int accumulate) {
  pdl_prologue();
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= C) return;  // warp-uniform
  double s, ss;
  finalize_sums(partials, P, C, c, s, ss);
  if ((threadIdx.x & 31) != 0) return;
  m1[c] = training ? (float)(s / count) : 0.f;
  m2[c] = training ? (float)(ss / count) : 0.f;
  if (dgamma != nullptr) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)ss;
  if (dbeta != nullptr) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s;
}

// y = relu(scale * z + shift)  -> fp32 and / or bf16
__global__ void bn_relu_apply_kernel(const float* __restrict__ z, const float* __restrict__ scale,
                                     const float* __restrict__ shift, float* __restrict__ y32,
                                     __nv_bfloat16* __restrict__ y16, long long n4, int C) {
  pdl_prologue();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i * 4) % C);
    const float4 v = reinterpret_cast<const float4*>(z)[i];
    const float4 s = *reinterpret_cast<const float4*>(scale + c);
    const float4 h = *reinterpret_cast<const float4*>(shift + c);
    float4 o;
    o.x = fmaxf(fmaf(s.x, v.x, h.x), 0.f);
    o.y = fmaxf(fmaf(s.y, v.y, h.y), 0.f);
    o.z = fmaxf(fmaf(s.z, v.z, h.z), 0.f);
    o.w = fmaxf(fmaf(s.w, v.w, h.w), 0.f);
    if (y32 != nullptr) reinterpret_cast<float4*>(y32)[i] = o;
    if (y16 != nullptr) {
      uint2 u;
      u.x = pack_bf16x2(o.x, o.y);
      u.y = pack_bf16x2(o.z, o.w);
      reinterpret_cast<uint2*>(y16)[i] = u;
    }
  }
}

// dz = scale * (dy - m1 - zhat * m2), dy = dout where scale * z + shift > 0   -> bf16 (next GEMM operand)
__global__ void bn_relu_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ z,
                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                         const float* __restrict__ m1, const float* __restrict__ m2,
                                         __nv_bfloat16* __restrict__ dz16, long long n4, int C) {
  pdl_prologue();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i * 4) % C);
    const float4 v = reinterpret_cast<const float4*>(z)[i];
    const float4 d = reinterpret_cast<const float4*>(dout)[i];
    const float vv[4] = {v.x, v.y, v.z, v.w}, dd[4] = {d.x, d.y, d.z, d.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float s = scale[c + j];
      const float dy = (fmaf(s, vv[j], shift[c + j]) > 0.f) ? dd[j] : 0.f;
      const float zh = (vv[j] - mean[c + j]) * rstd[c + j];
      o[j] = s * (dy - m1[c + j] - zh * m2[c + j]);
    }
    uint2 u;
    u.x = pack_bf16x2(o[0], o[1]);
    u.y = pack_bf16x2(o[2], o[3]);
    reinterpret_cast<uint2*>(dz16)[i] = u;
  }
}

// 3-NN inverse-distance interpolation (pointnet_util.py:401-408): w_j = (1 / (d_j + 1e-8)) / sum_j (1 / (d_j + 1e-8)).
__device__ __forceinline__ void three_nn_weights(const float* __restrict__ dist, long long row, int lane, long long i_in,
                                                 long long (&ij)[3], float (&wj)[3]) {
  float rcp = 0.f;
  if (lane < 3) rcp = 1.0f / (dist[row * 3 + lane] + 1e-8f);
  const float r0 = __shfl_sync(0xffffffffu, rcp, 0), r1 = __shfl_sync(0xffffffffu, rcp, 1),
              r2 = __shfl_sync(0xffffffffu, rcp, 2);
  const float norm = (r0 + r1) + r2;
  wj[0] = r0 / norm;
  wj[1] = r1 / norm;
  wj[2] = r2 / norm;
#pragma unroll
  for (int j = 0; j < 3; ++j) ij[j] = __shfl_sync(0xffffffffu, i_in, j);
}

__global__ void __launch_bounds__(256) three_nn_interp_fwd_kernel(const float* __restrict__ feats,
                                                                 const long long* __restrict__ idx,
                                                                 const float* __restrict__ dist,
                                                                 const float* __restrict__ addend,
                                                                 float* __restrict__ out, long long rows, int N, int S,
                                                                 int C) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const long long b = r / N;
    long long ij[3];
    float wj[3];
    three_nn_weights(dist, r, lane, lane < 3 ? idx[r * 3 + lane] : 0, ij, wj);
    const float* f0 = feats + ((size_t)b * S + (size_t)ij[0]) * C;
    const float* f1 = feats + ((size_t)b * S + (size_t)ij[1]) * C;
    const float* f2 = feats + ((size_t)b * S + (size_t)ij[2]) * C;
    for (int c = lane; c < C; c += 32) {
      float v = (f0[c] * wj[0] + f1[c] * wj[1]) + f2[c] * wj[2];
      if (addend != nullptr) v += addend[(size_t)r * C + c];
      out[(size_t)r * C + c] = v;
    }
  }
}

__global__ void __launch_bounds__(256) three_nn_interp_bwd_kernel(const float* __restrict__ dout,
                                                                 const long long* __restrict__ idx,
                                                                 const float* __restrict__ dist,
                                                                 float* __restrict__ dfeats, long long rows, int N,
                                                                 int S, int C) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const long long b = r / N;
    long long ij[3];
    float wj[3];
    three_nn_weights(dist, r, lane, lane < 3 ? idx[r * 3 + lane] : 0, ij, wj);
    for (int c = lane; c < C; c += 32) {
      const float d = dout[(size_t)r * C + c];
#pragma unroll
      for (int j = 0; j < 3; ++j) atomicAdd(dfeats + ((size_t)b * S + (size_t)ij[j]) * C + c, d * wj[j]);
    }
  }
}

// x fp32 [R, K] -> bf16 [R, 3K]: activation layout [hi | lo | hi], weight layout [hi | hi | lo] with hi = bf16(x),
// lo = bf16(x - hi). A GEMM over the tripled K then yields x_hi w_hi + x_lo w_hi + x_hi w_lo: fp32-grade products
// (error ~2^-17) from bf16 tensor-core operands, for the small per-point GEMMs that sit in front of a BatchNorm
// (where a bf16-rounded operand with |mean| >> std would be amplified by the normalisation).
__global__ void split_bf16x3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long R, int K,
                                    long long ldx, int weight_layout) {
  pdl_prologue();
  const long long n = R * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / K;
    const int k = (int)(i % K);
    const float v = x[r * ldx + k];
    const __nv_bfloat16 hi = __float2bfloat16(v);
    const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
    __nv_bfloat16* o = out + r * 3 * K;
    o[k] = hi;
    o[K + k] = weight_layout ? hi : lo;
    o[2 * K + k] = weight_layout ? lo : hi;
  }
}

inline int grid_for_rows(long long work_items, int C, int P) {
  (void)work_items;
  const int nchunks = (C + kChunk - 1) / kChunk;
  return nchunks * P;
}

inline int check_group(const SaGroup& a) {
  if (a.B <= 0 || a.N <= 0 || a.S <= 0 || a.K <= 0 || a.C1 <= 0 || a.ldw < 3) return S3D_ERR_BAD_SHAPE;
  if (a.K > 32 || (a.C1 & 3)) return S3D_ERR_UNSUPPORTED;
  if (a.uf == nullptr || a.xyz == nullptr || a.cxyz == nullptr || a.idx == nullptr || a.w1 == nullptr)
    return S3D_ERR_NULL;
  return S3D_OK;
}

}  // namespace

int sa_group_launch(int mode, const SaGroup& a, const SaExtra& e, int P, cudaStream_t stream) {
  const int rc = check_group(a);
  if (rc != S3D_OK) return rc;
  if (P <= 0) return S3D_ERR_BAD_SHAPE;
  const int grid = grid_for_rows((long long)a.B * a.S, a.C1, P);
  switch (mode) {
    case 0:
      if (e.partials == nullptr) return S3D_ERR_NULL;
      S3D_CUDA_OK(launch_pdl(sa_group_kernel<0>, dim3(grid), dim3(kWarps * 32), (size_t)(0), stream, a, e));
      break;
    case 1:
      if (e.scale == nullptr || e.shift == nullptr || e.a1_out == nullptr) return S3D_ERR_NULL;
      S3D_CUDA_OK(launch_pdl(sa_group_kernel<1>, dim3(grid), dim3(kWarps * 32), (size_t)(0), stream, a, e));
      break;
    case 2:
      if (e.mean == nullptr || e.rstd == nullptr || e.a1 == nullptr || e.da1 == nullptr || e.partials == nullptr)
        return S3D_ERR_NULL;
      S3D_CUDA_OK(launch_pdl(sa_group_kernel<2>, dim3(grid), dim3(kWarps * 32), (size_t)(0), stream, a, e));
      break;
    case 3:
      if (e.scale == nullptr || e.mean == nullptr || e.rstd == nullptr || e.m1 == nullptr || e.m2 == nullptr ||
          e.a1 == nullptr || e.da1 == nullptr || e.duf == nullptr || e.partials == nullptr)
        return S3D_ERR_NULL;
      S3D_CUDA_OK(launch_pdl(sa_group_kernel<3>, dim3(grid), dim3(kWarps * 32), (size_t)(0), stream, a, e));
      break;
    default: return S3D_ERR_UNSUPPORTED;
  }
  S3D_LAUNCH_OK();
  return S3D_OK;
}

}  // namespace s3d

// ------------------------------------------------------------------------------------------------------------------
// C ABI (declared in include/s3d_b200.h)
// ------------------------------------------------------------------------------------------------------------------
using namespace s3d;
static inline cudaStream_t st(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int s3d_sa_group_fwd_stats(const float* uf, const float* xyz, const float* cxyz, const int64_t* idx, const float* w1,
                           int ldw, int B, int N, int S, int K, int C1, float* partials, int P, void* stream) {
  SaGroup a{uf, xyz, cxyz, reinterpret_cast<const long long*>(idx), w1, ldw, B, N, S, K, C1};
  SaExtra e{};
  e.partials = partials;
  return sa_group_launch(0, a, e, P, st(stream));
}

int s3d_sa_group_fwd_act(const float* uf, const float* xyz, const float* cxyz, const int64_t* idx, const float* w1,
                         int ldw, int B, int N, int S, int K, int C1, const float* scale, const float* shift,
                         void* a1_bf16, int P, void* stream) {
  SaGroup a{uf, xyz, cxyz, reinterpret_cast<const long long*>(idx), w1, ldw, B, N, S, K, C1};
  SaExtra e{};
  e.scale = scale;
  e.shift = shift;
  e.a1_out = reinterpret_cast<__nv_bfloat16*>(a1_bf16);
  return sa_group_launch(1, a, e, P, st(stream));
}

int s3d_sa_group_bwd_stats(const float* uf, const float* xyz, const float* cxyz, const int64_t* idx, const float* w1,
                           int ldw, int B, int N, int S, int K, int C1, const float* mean, const float* rstd,
                           const void* a1_bf16, const void* da1_bf16, float* partials, int P, void* stream) {
  SaGroup a{uf, xyz, cxyz, reinterpret_cast<const long long*>(idx), w1, ldw, B, N, S, K, C1};
  SaExtra e{};
  e.mean = mean;
  e.rstd = rstd;
  e.a1 = reinterpret_cast<const __nv_bfloat16*>(a1_bf16);
  e.da1 = reinterpret_cast<const __nv_bfloat16*>(da1_bf16);
  e.partials = partials;
  return sa_group_launch(2, a, e, P, st(stream));
}

int s3d_sa_group_bwd_scatter(const float* uf, const float* xyz, const float* cxyz, const int64_t* idx, const float* w1,
                             int ldw, int B, int N, int S, int K, int C1, const float* scale, const float* mean,
                             const float* rstd, const float* m1, const float* m2, const void* a1_bf16,
                             const void* da1_bf16, float* duf, float* dwx_partials, int P, void* stream) {
  SaGroup a{uf, xyz, cxyz, reinterpret_cast<const long long*>(idx), w1, ldw, B, N, S, K, C1};
  SaExtra e{};
  e.scale = scale;
  e.mean = mean;
  e.rstd = rstd;
  e.m1 = m1;
  e.m2 = m2;
  e.a1 = reinterpret_cast<const __nv_bfloat16*>(a1_bf16);
  e.da1 = reinterpret_cast<const __nv_bfloat16*>(da1_bf16);
  e.duf = duf;
  e.partials = dwx_partials;
  return sa_group_launch(3, a, e, P, st(stream));
}

int s3d_sa_group_reduce(const float* z2, int64_t G, int K, int C, float* zmax, float* zmin, uint8_t* kmax,
                        uint8_t* kmin, float* partials, int P, void* stream) {
  if (G <= 0 || K <= 0 || C <= 0 || P <= 0) return S3D_ERR_BAD_SHAPE;
  if (K > 255 || (C & 3)) return S3D_ERR_UNSUPPORTED;
  if (z2 == nullptr || zmax == nullptr || zmin == nullptr || kmax == nullptr || kmin == nullptr || partials == nullptr)
    return S3D_ERR_NULL;
  S3D_CUDA_OK(launch_pdl(sa_group_reduce_kernel, dim3(grid_for_rows(G, C, P)), dim3(kWarps * 32), (size_t)(0), st(stream), z2, G, K, C, zmax, zmin, kmax, kmin,
                                                                                  partials));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int s3d_sa_pool_select(const float* zmax, const float* zmin, const uint8_t* kmax, const uint8_t* kmin,
                       const float* scale, const float* shift, float* out, float* zsel, uint8_t* ksel, int64_t G, int C,
                       void* stream) {
  if (G <= 0 || C <= 0) return S3D_ERR_BAD_SHAPE;
  if (zmax == nullptr || zmin == nullptr || kmax == nullptr || kmin == nullptr || scale == nullptr ||
      shift == nullptr || out == nullptr || zsel == nullptr || ksel == nullptr)
    return S3D_ERR_NULL;
  const long long n = (long long)G * C;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  S3D_CUDA_OK(launch_pdl(sa_pool_select_kernel, dim3((int)blocks), dim3(256), (size_t)(0), st(stream), zmax, zmin, kmax, kmin, scale, shift, out, zsel, ksel, n, C));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int s3d_sa_dz2_expand(const float* z2, const float* dout, const float* zsel, const uint8_t* ksel, const float* scale,
                      const float* shift, const float* mean, const float* rstd, const float* m1, const float* m2,
                      void* dz2_bf16, int64_t G, int K, int C, int P, void* stream) {
  if (G <= 0 || K <= 0 || C <= 0 || P <= 0) return S3D_ERR_BAD_SHAPE;
  if (K > 255 || (C & 3)) return S3D_ERR_UNSUPPORTED;
  if (z2 == nullptr || dout == nullptr || zsel == nullptr || ksel == nullptr || scale == nullptr || shift == nullptr ||
      mean == nullptr || rstd == nullptr || m1 == nullptr || m2 == nullptr || dz2_bf16 == nullptr)
    return S3D_ERR_NULL;
  S3D_CUDA_OK(launch_pdl(sa_dz2_expand_kernel, dim3(grid_for_rows(G, C, P)), dim3(kWarps * 32), (size_t)(0), st(stream), 
      z2, dout, zsel, ksel, scale, shift, mean, rstd, m1, m2, reinterpret_cast<__nv_bfloat16*>(dz2_bf16), G, K, C));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int s3d_bn_rows_stats(const float* z, int64_t R, int C, float* partials, int P, void* stream) {
  if (R <= 0 || C <= 0 || P <= 0) return S3D_ERR_BAD_SHAPE;
  if (C & 3) return S3D_ERR_UNSUPPORTED;
  if (z == nullptr || partials == nullptr) return S3D_ERR_NULL;
  S3D_CUDA_OK(launch_pdl(bn_rows_stats_kernel<0>, dim3(grid_for_rows(R, C, P)), dim3(kWarps * 32), (size_t)(0), st(stream), z, nullptr, nullptr, nullptr, nullptr,
                                                                                   nullptr, R, C, partials));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int s3d_bn_rows_bwd_stats(const float* dout, const float* z, const float* scale, const float* shift, const float* mean,
                          const float* rstd, int64_t R, int C, float* partials, int P, void* stream) {
  if (R <= 0 || C <= 0 || P <= 0) return S3D_ERR_BAD_SHAPE;
  if (C & 3) return S3D_ERR_UNSUPPORTED;
  if (dout == nullptr || z == nullptr || scale == nullptr || shift == nullptr || mean == nullptr || rstd == nullptr ||
      partials == nullptr)
    return S3D_ERR_NULL;
  S3D_CUDA_OK(launch_pdl(bn_rows_stats_kernel<1>, dim3(grid_for_rows(R, C, P)), dim3(kWarps * 32), (size_t)(0), st(stream), z, dout, scale, shift, mean, rstd, R,
                                                                                   C, partials));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int s3d_bn_finalize_fwd(const float* partials, int P, int C, double count, const float* gamma, const float* beta,
                        float eps, float momentum, float* running_mean, float* running_var, float* mean, float* rstd,
                        float* scale, float* shift, void* stream) {
  if (P <= 0 || C <= 0 || count <= 0.0) return S3D_ERR_BAD_SHAPE;
  if (partials == nullptr || mean == nullptr || rstd == nullptr || scale == nullptr || shift == nullptr)
    return S3D_ERR_NULL;
  S3D_CUDA_OK(launch_pdl(bn_finalize_fwd_kernel, dim3((C + 3) / 4), dim3(128), (size_t)(0), st(stream), partials, P, C, count, gamma, beta, eps, momentum,
                                                                   running_mean, running_var, mean, rstd, scale, shift));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int s3d_bn_finalize_bwd(const float* partials, int P, int C, double count, int training, float* m1, float* m2,
                        float* dgamma, float* dbeta, int accumulate, void* stream) {
  if (P <= 0 || C <= 0 || count <= 0.0) return S3D_ERR_BAD_SHAPE;
  if (partials == nullptr || m1 == nullptr || m2 == nullptr) return S3D_ERR_NULL;
  S3D_CUDA_OK(launch_pdl(bn_finalize_bwd_kernel, dim3((C + 3) / 4), dim3(128), (size_t)(0), st(stream), partials, P, C, count, training, m1, m2, dgamma, dbeta,
                                                                   accumulate));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int s3d_bn_relu_apply(const float* z, const float* scale, const float* shift, float* y_f32, void* y_bf16, int64_t R,
                      int C, void* stream) {
  if (R <= 0 || C <= 0) return S3D_ERR_BAD_SHAPE;
  if (C % 4 != 0) return S3D_ERR_UNSUPPORTED;
  if (z == nullptr || scale == nullptr || shift == nullptr || (y_f32 == nullptr && y_bf16 == nullptr))
    return S3D_ERR_NULL;
  const long long n4 = (long long)R * C / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  S3D_CUDA_OK(launch_pdl(bn_relu_apply_kernel, dim3((int)blocks), dim3(256), (size_t)(0), st(stream), z, scale, shift, y_f32,
                                                            reinterpret_cast<__nv_bfloat16*>(y_bf16), n4, C));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int s3d_bn_relu_bwd_apply(const float* dout, const float* z, const float* scale, const float* shift, const float* mean,
                          const float* rstd, const float* m1, const float* m2, void* dz_bf16, int64_t R, int C,
                          void* stream) {
  if (R <= 0 || C <= 0) return S3D_ERR_BAD_SHAPE;
  if (C % 4 != 0) return S3D_ERR_UNSUPPORTED;
  if (dout == nullptr || z == nullptr || scale == nullptr || shift == nullptr || mean == nullptr || rstd == nullptr ||
      m1 == nullptr || m2 == nullptr || dz_bf16 == nullptr)
    return S3D_ERR_NULL;
  const long long n4 = (long long)R * C / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  S3D_CUDA_OK(launch_pdl(bn_relu_bwd_apply_kernel, dim3((int)blocks), dim3(256), (size_t)(0), st(stream), dout, z, scale, shift, mean, rstd, m1, m2,
                                                                reinterpret_cast<__nv_bfloat16*>(dz_bf16), n4, C));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int s3d_split_bf16x3(const float* x, void* out_bf16, int64_t R, int K, int64_t ldx, int weight_layout, void* stream) {
  if (R <= 0 || K <= 0 || ldx < K) return S3D_ERR_BAD_SHAPE;
  if (x == nullptr || out_bf16 == nullptr) return S3D_ERR_NULL;
  const long long n = (long long)R * K;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  S3D_CUDA_OK(launch_pdl(split_bf16x3_kernel, dim3((int)blocks), dim3(256), (size_t)(0), st(stream), x, reinterpret_cast<__nv_bfloat16*>(out_bf16), R, K, ldx,
                                                           weight_layout));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int s3d_three_nn_interp_fwd(const float* feats, const int64_t* idx, const float* dist, const float* addend, float* out,
                            int B, int S, int N, int C, void* stream) {
  if (B <= 0 || S <= 0 || N <= 0 || C <= 0) return S3D_ERR_BAD_SHAPE;
  if (feats == nullptr || idx == nullptr || dist == nullptr || out == nullptr) return S3D_ERR_NULL;
  const long long rows = (long long)B * N;
  long long blocks = (rows + 7) / 8;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  S3D_CUDA_OK(launch_pdl(three_nn_interp_fwd_kernel, dim3((int)blocks), dim3(256), (size_t)(0), st(stream), feats, reinterpret_cast<const long long*>(idx), dist,
                                                                  addend, out, rows, N, S, C));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int s3d_three_nn_interp_bwd(const float* dout, const int64_t* idx, const float* dist, float* dfeats, int B, int S, int N,
                            int C, void* stream) {
  if (B <= 0 || S <= 0 || N <= 0 || C <= 0) return S3D_ERR_BAD_SHAPE;
  if (dout == nullptr || idx == nullptr || dist == nullptr || dfeats == nullptr) return S3D_ERR_NULL;
  S3D_CUDA_OK(cudaMemsetAsync(dfeats, 0, sizeof(float) * (size_t)B * S * C, st(stream)));
  const long long rows = (long long)B * N;
  long long blocks = (rows + 7) / 8;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  S3D_CUDA_OK(launch_pdl(three_nn_interp_bwd_kernel, dim3((int)blocks), dim3(256), (size_t)(0), st(stream), dout, reinterpret_cast<const long long*>(idx), dist,
                                                                  dfeats, rows, N, S, C));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

}  // extern "C"
