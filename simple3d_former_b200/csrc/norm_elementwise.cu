// HBM-bound kernels of the encoder hot path: LayerNorm forward/backward (timm Block.norm1/norm2, reference
// vit_3d_2d_pretrain.py:287 eps=1e-6; nn.TransformerEncoderLayer norms eps=1e-5), fp32->bf16 weight shadowing,
// bias-gradient column sums, voxel patch gather (Conv3d k=s=cell as a GEMM operand, embed_layer_3d_modality.py:22-24),
// and the fused Adam step (train_cls_voxel.py:195). All are one-pass, 128-bit vectorised, grid sized to the SM count.
#include <cstdlib>

#include "kernels.h"

namespace s3d {

// ------------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row, the row lives in registers (D <= 1024, D % 4 == 0).
// x fp32 [T,D] -> y (bf16 and/or fp32), mean/rstd fp32 [T]. Optional fused "x = a + b" prologue (post-norm layers).
// ------------------------------------------------------------------------------------------------
template <int VEC_ITERS>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ addend,
                                                           float* __restrict__ sum_out,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_f32,
                                                           float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                           int T, int D, float eps) {
  pdl_prologue();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nvec = D >> 2;
  for (int row = warp; row < T; row += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
    float4 v[VEC_ITERS];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VEC_ITERS; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        v[i] = xr[c];
        if (addend != nullptr) {
          const float4 a = reinterpret_cast<const float4*>(addend + (size_t)row * D)[c];
          v[i].x += a.x; v[i].y += a.y; v[i].z += a.z; v[i].w += a.w;
          if (sum_out != nullptr) reinterpret_cast<float4*>(sum_out + (size_t)row * D)[c] = v[i];
        }
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      } else {
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    s = warp_sum(s);
    const float mean = s / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VEC_ITERS; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    q = warp_sum(q);
    const float rstd = rsqrtf(q / (float)D + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VEC_ITERS; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 g = reinterpret_cast<const float4*>(gamma)[c];
        const float4 b = reinterpret_cast<const float4*>(beta)[c];
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + b.x;
        o.y = (v[i].y - mean) * rstd * g.y + b.y;
        o.z = (v[i].z - mean) * rstd * g.z + b.z;
        o.w = (v[i].w - mean) * rstd * g.w + b.w;
        if (y_bf16 != nullptr) {
          uint2 u;
          u.x = pack_bf16x2(o.x, o.y);
          u.y = pack_bf16x2(o.z, o.w);
          reinterpret_cast<uint2*>(y_bf16 + (size_t)row * D)[c] = u;
        }
        if (y_f32 != nullptr) reinterpret_cast<float4*>(y_f32 + (size_t)row * D)[c] = o;
      }
    }
  }
}

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// LayerNorm forward, bulk-async pipelined variant (large T, 16-byte aligned rows): each warp streams its rows through a
// ring of shared-memory slots with cp.async.bulk + mbarrier (see the backward variant below for the rationale).
// ------------------------------------------------------------------------------------------------
constexpr int kLnFwdWarps = 16;

template <int VEC_ITERS, int STAGES>
__global__ void __launch_bounds__(kLnFwdWarps * 32, 1)
    layernorm_fwd_pipe_kernel(const float* __restrict__ x, const float* __restrict__ addend, float* __restrict__ sum_out,
                              const float* __restrict__ gamma, const float* __restrict__ beta,
                              __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_f32,
                              float* __restrict__ mean_out, float* __restrict__ rstd_out, int T, int D, float eps) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t lnf_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nvec = D >> 2;
  const uint32_t x_bytes = (uint32_t)D * 4u;
  const uint32_t row_bytes = addend != nullptr ? 2u * x_bytes : x_bytes;
  uint8_t* ring = lnf_smem + (size_t)wib * STAGES * row_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(lnf_smem + (size_t)kLnFwdWarps * STAGES * row_bytes) + wib * STAGES;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
  }
  fence_barrier_init();
  __syncthreads();
  const long long warp = (long long)blockIdx.x * kLnFwdWarps + wib;
  const long long nwarps = (long long)gridDim.x * kLnFwdWarps;
  auto issue = [&](long long row, int s) {  // lane 0 only
    uint8_t* dst = ring + (size_t)s * row_bytes;
    mbar_expect_tx(&bars[s], row_bytes);
    bulk_g2s(dst, x + (size_t)row * D, x_bytes, &bars[s]);
    if (addend != nullptr) bulk_g2s(dst + x_bytes, addend + (size_t)row * D, x_bytes, &bars[s]);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      const long long row = warp + (long long)s * nwarps;
      if (row < T) issue(row, s);
    }
  }
  float4 gm[VEC_ITERS], bt[VEC_ITERS];
#pragma unroll
  for (int i = 0; i < VEC_ITERS; ++i) {
    const int c = lane + 32 * i;
    gm[i] = c < nvec ? reinterpret_cast<const float4*>(gamma)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    bt[i] = c < nvec ? reinterpret_cast<const float4*>(beta)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int it = 0;
  for (long long row = warp; row < T; row += nwarps, ++it) {
    const int s = it % STAGES;
    const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
    const uint8_t* buf = ring + (size_t)s * row_bytes;
    mbar_wait(&bars[s], parity);
    float4 v[VEC_ITERS];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VEC_ITERS; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        v[i] = reinterpret_cast<const float4*>(buf)[c];
        if (addend != nullptr) {
          const float4 a = reinterpret_cast<const float4*>(buf + x_bytes)[c];
          v[i].x += a.x; v[i].y += a.y; v[i].z += a.z; v[i].w += a.w;
          if (sum_out != nullptr) reinterpret_cast<float4*>(sum_out + (size_t)row * D)[c] = v[i];
        }
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      } else {
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // slot s is in registers now: refill it for the row STAGES ahead before doing the arithmetic
    __syncwarp();
    const long long next = row + (long long)STAGES * nwarps;
    if (lane == 0 && next < T) issue(next, s);
    sum = warp_sum(sum);
    const float mean = sum / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VEC_ITERS; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    q = warp_sum(q);
    const float rstd = rsqrtf(q / (float)D + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VEC_ITERS; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        float4 o;
        o.x = (v[i].x - mean) * rstd * gm[i].x + bt[i].x;
        o.y = (v[i].y - mean) * rstd * gm[i].y + bt[i].y;
        o.z = (v[i].z - mean) * rstd * gm[i].z + bt[i].z;
        o.w = (v[i].w - mean) * rstd * gm[i].w + bt[i].w;
        if (y_bf16 != nullptr) {
          uint2 u;
          u.x = pack_bf16x2(o.x, o.y);
          u.y = pack_bf16x2(o.z, o.w);
          reinterpret_cast<uint2*>(y_bf16 + (size_t)row * D)[c] = u;
        }
        if (y_f32 != nullptr) reinterpret_cast<float4*>(y_f32 + (size_t)row * D)[c] = o;
      }
    }
  }
}

template <int I>
static int launch_ln_fwd_pipe(const float* x, const float* addend, float* sum_out, const float* gamma, const float* beta,
                              __nv_bfloat16* yb, float* y_f32, float* mean, float* rstd, int T, int D, float eps,
                              cudaStream_t stream) {
  const size_t row_bytes = (size_t)D * 4 * (addend != nullptr ? 2 : 1);
  const int stages = addend != nullptr ? 2 : 3;
  const size_t shmem = (size_t)kLnFwdWarps * stages * (row_bytes + sizeof(uint64_t));
  if (shmem > 227 * 1024) return S3D_ERR_UNSUPPORTED;
  long long blocks = ((long long)T + kLnFwdWarps - 1) / kLnFwdWarps;
  if (blocks > num_sms()) blocks = num_sms();
  if (stages == 3) {
    auto kern = layernorm_fwd_pipe_kernel<I, 3>;
    S3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    S3D_CUDA_OK(launch_pdl(kern, dim3((int)blocks), dim3(kLnFwdWarps * 32), (size_t)(shmem), stream, x, addend, sum_out, gamma, beta, yb, y_f32, mean, rstd, T, D,
                                                           eps));
  } else {
    auto kern = layernorm_fwd_pipe_kernel<I, 2>;
    S3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    S3D_CUDA_OK(launch_pdl(kern, dim3((int)blocks), dim3(kLnFwdWarps * 32), (size_t)(shmem), stream, x, addend, sum_out, gamma, beta, yb, y_f32, mean, rstd, T, D,
                                                           eps));
  }
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int layernorm_fwd(const float* x, const float* addend, float* sum_out, const float* gamma, const float* beta,
                  void* y_bf16, float* y_f32, float* mean, float* rstd, int T, int D, float eps, cudaStream_t stream) {
  if (T <= 0 || D <= 0 || D % 4 != 0 || D > 1024) return S3D_ERR_BAD_SHAPE;
  if (x == nullptr || gamma == nullptr || beta == nullptr) return S3D_ERR_NULL;
  const int warps_per_block = 8;
  long long blocks = ((long long)T + warps_per_block - 1) / warps_per_block;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  const int iters = (D / 4 + 31) / 32;
  auto yb = reinterpret_cast<__nv_bfloat16*>(y_bf16);
  static const bool no_pipe = getenv("S3D_LN_FWD_NO_PIPE") != nullptr;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(addend)) & 15) == 0;
  if (!no_pipe && aligned && (long long)T >= (long long)num_sms() * kLnFwdWarps * 4) {
    int rc = S3D_ERR_UNSUPPORTED;
#define S3D_LN_FWD_PIPE(I) \
  rc = launch_ln_fwd_pipe<I>(x, addend, sum_out, gamma, beta, yb, y_f32, mean, rstd, T, D, eps, stream)
    switch (iters) {
      case 1: S3D_LN_FWD_PIPE(1); break;
      case 2: S3D_LN_FWD_PIPE(2); break;
      case 3: S3D_LN_FWD_PIPE(3); break;
      case 4: S3D_LN_FWD_PIPE(4); break;
      case 5: S3D_LN_FWD_PIPE(5); break;
      case 6: S3D_LN_FWD_PIPE(6); break;
      case 7: S3D_LN_FWD_PIPE(7); break;
      default: S3D_LN_FWD_PIPE(8); break;
    }
#undef S3D_LN_FWD_PIPE
    if (rc != S3D_ERR_UNSUPPORTED) return rc;
  }
#define S3D_LN_FWD(I)                                                                                              \
  S3D_CUDA_OK(launch_pdl(layernorm_fwd_kernel<I>, dim3((int)blocks), dim3(warps_per_block * 32), (size_t)(0), stream, x, addend, sum_out, gamma, beta, yb, \
                                                                            y_f32, mean, rstd, T, D, eps))
  switch (iters) {
    case 1: S3D_LN_FWD(1); break;
    case 2: S3D_LN_FWD(2); break;
    case 3: S3D_LN_FWD(3); break;
    case 4: S3D_LN_FWD(4); break;
    case 5: S3D_LN_FWD(5); break;
    case 6: S3D_LN_FWD(6); break;
    case 7: S3D_LN_FWD(7); break;
    default: S3D_LN_FWD(8); break;
  }
#undef S3D_LN_FWD
  S3D_LAUNCH_OK();
  return S3D_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward. dy (bf16 or fp32) is the gradient w.r.t. the LN output; `dres` (optional, fp32) is the
// gradient arriving through the residual branch and is added to dx. Writes dx fp32 and (optionally) a bf16 copy
// that feeds the next tensor-core GEMM. dgamma/dbeta are accumulated with one atomicAdd per CTA per column.
// ------------------------------------------------------------------------------------------------
template <int VEC_ITERS, bool DY_BF16>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const void* __restrict__ dy_, const float* __restrict__ x,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ mean_in,
                                                           const float* __restrict__ rstd_in,
                                                           const float* __restrict__ dres, float* __restrict__ dx,
                                                           __nv_bfloat16* __restrict__ dx_bf16,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                           float* __restrict__ dxsum, int T, int D) {
  pdl_prologue();
  extern __shared__ float red[];  // [3][D]
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nvec = D >> 2;
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) red[i] = 0.f;
  __syncthreads();

  float4 acc_g[VEC_ITERS], acc_b[VEC_ITERS], acc_x[VEC_ITERS];
#pragma unroll
  for (int i = 0; i < VEC_ITERS; ++i) {
    acc_g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc_b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc_x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int row = warp; row < T; row += nwarps) {
    const float mean = mean_in[row];
    const float rstd = rstd_in[row];
    float4 xh[VEC_ITERS], gy[VEC_ITERS];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VEC_ITERS; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 xv = reinterpret_cast<const float4*>(x + (size_t)row * D)[c];
        float4 dyv;
        if (DY_BF16) {
          const uint2 u = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy_) + (size_t)row * D)[c];
          const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
          dyv = make_float4(a.x, a.y, b.x, b.y);
        } else {
          dyv = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + (size_t)row * D)[c];
        }
        const float4 g = reinterpret_cast<const float4*>(gamma)[c];
        xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
        gy[i] = make_float4(dyv.x * g.x, dyv.y * g.y, dyv.z * g.z, dyv.w * g.w);
        s1 += (gy[i].x + gy[i].y) + (gy[i].z + gy[i].w);
        s2 += (gy[i].x * xh[i].x + gy[i].y * xh[i].y) + (gy[i].z * xh[i].z + gy[i].w * xh[i].w);
        acc_g[i].x += dyv.x * xh[i].x; acc_g[i].y += dyv.y * xh[i].y;
        acc_g[i].z += dyv.z * xh[i].z; acc_g[i].w += dyv.w * xh[i].w;
        acc_b[i].x += dyv.x; acc_b[i].y += dyv.y; acc_b[i].z += dyv.z; acc_b[i].w += dyv.w;
      }
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
#pragma unroll
    for (int i = 0; i < VEC_ITERS; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        float4 o;
        o.x = rstd * (gy[i].x - s1 - xh[i].x * s2);
        o.y = rstd * (gy[i].y - s1 - xh[i].y * s2);
        o.z = rstd * (gy[i].z - s1 - xh[i].z * s2);
        o.w = rstd * (gy[i].w - s1 - xh[i].w * s2);
        if (dres != nullptr) {
          const float4 r = reinterpret_cast<const float4*>(dres + (size_t)row * D)[c];
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        reinterpret_cast<float4*>(dx + (size_t)row * D)[c] = o;
        acc_x[i].x += o.x; acc_x[i].y += o.y; acc_x[i].z += o.z; acc_x[i].w += o.w;
        if (dx_bf16 != nullptr) {
          uint2 u;
          u.x = pack_bf16x2(o.x, o.y);
          u.y = pack_bf16x2(o.z, o.w);
          reinterpret_cast<uint2*>(dx_bf16 + (size_t)row * D)[c] = u;
        }
      }
    }
  }
  // CTA-level reduction of the parameter gradients, then one atomic per column per CTA.
  if (dgamma != nullptr) {
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      if (w == warp_in_block) {
#pragma unroll
        for (int i = 0; i < VEC_ITERS; ++i) {
          const int c = lane + 32 * i;
          if (c < nvec) {
            float* rg = red + 4 * c;
            float* rb = red + D + 4 * c;
            rg[0] += acc_g[i].x; rg[1] += acc_g[i].y; rg[2] += acc_g[i].z; rg[3] += acc_g[i].w;
            rb[0] += acc_b[i].x; rb[1] += acc_b[i].y; rb[2] += acc_b[i].z; rb[3] += acc_b[i].w;
            if (dxsum != nullptr) {
              float* rx = red + 2 * D + 4 * c;
              rx[0] += acc_x[i].x; rx[1] += acc_x[i].y; rx[2] += acc_x[i].z; rx[3] += acc_x[i].w;
            }
          }
        }
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      atomicAdd(dgamma + i, red[i]);
      atomicAdd(dbeta + i, red[D + i]);
      if (dxsum != nullptr) atomicAdd(dxsum + i, red[2 * D + i]);  // column sums of dx: the bias gradient of the Linear that
    }                                                                // produced this LayerNorm's input's other branch
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward, bulk-async pipelined variant (D % 8 == 0, the shapes of the encoder blocks).
// The register-resident kernel above keeps only ~16 warps per SM busy (128 registers per thread) and each warp
// alternates between a load phase and a compute/store phase, so the memory system sees ~40 % duty (ncu: 3.4 TB/s).
// Here every warp owns a ring of kLnStages row buffers in shared memory; lane 0 streams the rows (x, dy, dres) in with
// cp.async.bulk + an mbarrier per slot (complete_tx), always kLnStages rows ahead of the arithmetic, so the loads of the
// next rows are in flight while the current row is reduced, and no registers are spent on data that is not yet used.
// ------------------------------------------------------------------------------------------------
constexpr int kLnWarps = 12;
constexpr int kLnStages = 2;

template <int VEC_ITERS, bool DY_BF16>
__global__ void __launch_bounds__(kLnWarps * 32, 1)
    layernorm_bwd_pipe_kernel(const void* __restrict__ dy_, const float* __restrict__ x, const float* __restrict__ gamma,
                              const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                              const float* __restrict__ dres, float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_bf16,
                              float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dxsum, int T,
                              int D) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int nvec = D >> 2;
  const uint32_t x_bytes = (uint32_t)D * 4u;
  const uint32_t dy_bytes = DY_BF16 ? (uint32_t)D * 2u : (uint32_t)D * 4u;
  const uint32_t dres_bytes = dres != nullptr ? (uint32_t)D * 4u : 0u;
  const uint32_t row_bytes = x_bytes + dy_bytes + dres_bytes;
  // layout: [kLnWarps][kLnStages][row_bytes] | red [3][D] f32 | mbarriers [kLnWarps][kLnStages]
  uint8_t* ring = ln_smem + (size_t)wib * kLnStages * row_bytes;
  float* red = reinterpret_cast<float*>(ln_smem + (size_t)kLnWarps * kLnStages * row_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 3 * D) + wib * kLnStages;
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) red[i] = 0.f;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kLnStages; ++s) mbar_init(&bars[s], 1);
  }
  fence_barrier_init();
  __syncthreads();

  const long long warp = (long long)blockIdx.x * kLnWarps + wib;
  const long long nwarps = (long long)gridDim.x * kLnWarps;
  auto issue = [&](long long row, int s) {  // lane 0 only
    uint8_t* dst = ring + (size_t)s * row_bytes;
    mbar_expect_tx(&bars[s], row_bytes);
    bulk_g2s(dst, x + (size_t)row * D, x_bytes, &bars[s]);
    bulk_g2s(dst + x_bytes, reinterpret_cast<const uint8_t*>(dy_) + (size_t)row * dy_bytes, dy_bytes, &bars[s]);
    if (dres_bytes) bulk_g2s(dst + x_bytes + dy_bytes, dres + (size_t)row * D, dres_bytes, &bars[s]);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kLnStages; ++s) {
      const long long row = warp + (long long)s * nwarps;
      if (row < T) issue(row, s);
    }
  }
  float4 gm[VEC_ITERS];
  float4 acc_g[VEC_ITERS], acc_b[VEC_ITERS], acc_x[VEC_ITERS];
#pragma unroll
  for (int i = 0; i < VEC_ITERS; ++i) {
    const int c = lane + 32 * i;
    gm[i] = c < nvec ? reinterpret_cast<const float4*>(gamma)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    acc_g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc_b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc_x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int it = 0;
  for (long long row = warp; row < T; row += nwarps, ++it) {
    const int s = it % kLnStages;
    const uint32_t parity = (uint32_t)(it / kLnStages) & 1u;
    const float mean = mean_in[row];
    const float rstd = rstd_in[row];
    const uint8_t* buf = ring + (size_t)s * row_bytes;
    mbar_wait(&bars[s], parity);
    float4 xh[VEC_ITERS], gy[VEC_ITERS];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VEC_ITERS; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 xv = reinterpret_cast<const float4*>(buf)[c];
        float4 dyv;
        if (DY_BF16) {
          const uint2 u = reinterpret_cast<const uint2*>(buf + x_bytes)[c];
          const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
          dyv = make_float4(a.x, a.y, b.x, b.y);
        } else {
          dyv = reinterpret_cast<const float4*>(buf + x_bytes)[c];
        }
        xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
        gy[i] = make_float4(dyv.x * gm[i].x, dyv.y * gm[i].y, dyv.z * gm[i].z, dyv.w * gm[i].w);
        s1 += (gy[i].x + gy[i].y) + (gy[i].z + gy[i].w);
        s2 += (gy[i].x * xh[i].x + gy[i].y * xh[i].y) + (gy[i].z * xh[i].z + gy[i].w * xh[i].w);
        acc_g[i].x += dyv.x * xh[i].x; acc_g[i].y += dyv.y * xh[i].y;
        acc_g[i].z += dyv.z * xh[i].z; acc_g[i].w += dyv.w * xh[i].w;
        acc_b[i].x += dyv.x; acc_b[i].y += dyv.y; acc_b[i].z += dyv.z; acc_b[i].w += dyv.w;
      }
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
#pragma unroll
    for (int i = 0; i < VEC_ITERS; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        float4 o;
        o.x = rstd * (gy[i].x - s1 - xh[i].x * s2);
        o.y = rstd * (gy[i].y - s1 - xh[i].y * s2);
        o.z = rstd * (gy[i].z - s1 - xh[i].z * s2);
        o.w = rstd * (gy[i].w - s1 - xh[i].w * s2);
        if (dres_bytes) {
          const float4 r = reinterpret_cast<const float4*>(buf + x_bytes + dy_bytes)[c];
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        reinterpret_cast<float4*>(dx + (size_t)row * D)[c] = o;
        acc_x[i].x += o.x; acc_x[i].y += o.y; acc_x[i].z += o.z; acc_x[i].w += o.w;
        if (dx_bf16 != nullptr) {
          uint2 u;
          u.x = pack_bf16x2(o.x, o.y);
          u.y = pack_bf16x2(o.z, o.w);
          reinterpret_cast<uint2*>(dx_bf16 + (size_t)row * D)[c] = u;
        }
      }
    }
    // every lane has finished reading slot s: hand it back to the copy engine for the row kLnStages ahead
    __syncwarp();
    const long long next = row + (long long)kLnStages * nwarps;
    if (lane == 0 && next < T) issue(next, s);
  }
  if (dgamma != nullptr) {
    for (int w = 0; w < kLnWarps; ++w) {
      if (w == wib) {
#pragma unroll
        for (int i = 0; i < VEC_ITERS; ++i) {
          const int c = lane + 32 * i;
          if (c < nvec) {
            float* rg = red + 4 * c;
            float* rb = red + D + 4 * c;
            rg[0] += acc_g[i].x; rg[1] += acc_g[i].y; rg[2] += acc_g[i].z; rg[3] += acc_g[i].w;
            rb[0] += acc_b[i].x; rb[1] += acc_b[i].y; rb[2] += acc_b[i].z; rb[3] += acc_b[i].w;
            if (dxsum != nullptr) {
              float* rx = red + 2 * D + 4 * c;
              rx[0] += acc_x[i].x; rx[1] += acc_x[i].y; rx[2] += acc_x[i].z; rx[3] += acc_x[i].w;
            }
          }
        }
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      atomicAdd(dgamma + i, red[i]);
      atomicAdd(dbeta + i, red[D + i]);
      if (dxsum != nullptr) atomicAdd(dxsum + i, red[2 * D + i]);  // column sums of dx: the bias gradient of the Linear that
    }                                                                // produced this LayerNorm's input's other branch
  }
}

template <int I, bool BF>
static int launch_ln_bwd_pipe(const void* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                              const float* dres, float* dx, __nv_bfloat16* dxb, float* dgamma, float* dbeta, float* dxsum,
                              int T, int D, cudaStream_t stream) {
  const size_t row_bytes = (size_t)D * 4 + (BF ? (size_t)D * 2 : (size_t)D * 4) + (dres != nullptr ? (size_t)D * 4 : 0);
  const size_t shmem = (size_t)kLnWarps * kLnStages * row_bytes + 3 * (size_t)D * sizeof(float) +
                       (size_t)kLnWarps * kLnStages * sizeof(uint64_t);
  if (shmem > 227 * 1024) return S3D_ERR_UNSUPPORTED;
  auto kern = layernorm_bwd_pipe_kernel<I, BF>;
  S3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
  long long blocks = ((long long)T + kLnWarps - 1) / kLnWarps;
  if (blocks > num_sms()) blocks = num_sms();
  S3D_CUDA_OK(launch_pdl(kern, dim3((int)blocks), dim3(kLnWarps * 32), (size_t)(shmem), stream, dy, x, gamma, mean, rstd, dres, dx, dxb, dgamma, dbeta, dxsum, T, D));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int layernorm_bwd(const void* dy, int dy_is_bf16, const float* x, const float* gamma, const float* mean,
                  const float* rstd, const float* dres, float* dx, void* dx_bf16, float* dgamma, float* dbeta,
                  float* dxsum, int T, int D, cudaStream_t stream) {
  if (T <= 0 || D <= 0 || D % 4 != 0 || D > 1024) return S3D_ERR_BAD_SHAPE;
  if (dxsum != nullptr && dgamma == nullptr) return S3D_ERR_NULL;  // the column sums share the parameter-gradient reduction
  if (dy == nullptr || x == nullptr || gamma == nullptr || mean == nullptr || rstd == nullptr || dx == nullptr)
    return S3D_ERR_NULL;
  const int warps_per_block = 8;
  long long blocks = ((long long)T + warps_per_block - 1) / warps_per_block;
  const long long cap = (long long)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  const int iters = (D / 4 + 31) / 32;
  const size_t shmem = 3 * (size_t)D * sizeof(float);
  auto dxb = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
  // large problems with 16-byte-aligned rows: bulk-async pipelined variant (rows streamed through shared memory)
  static const bool no_pipe = getenv("S3D_LN_BWD_NO_PIPE") != nullptr;
  const bool aligned = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) |
                         reinterpret_cast<uintptr_t>(dres)) & 15) == 0;
  if (!no_pipe && D % 8 == 0 && aligned && (long long)T >= (long long)num_sms() * kLnWarps * 4 && iters <= 8) {
    int rc = S3D_ERR_UNSUPPORTED;
#define S3D_LN_PIPE(I)                                                                                            \
  rc = dy_is_bf16 ? launch_ln_bwd_pipe<I, true>(dy, x, gamma, mean, rstd, dres, dx, dxb, dgamma, dbeta, dxsum, T, D, stream) \
                  : launch_ln_bwd_pipe<I, false>(dy, x, gamma, mean, rstd, dres, dx, dxb, dgamma, dbeta, dxsum, T, D, stream)
    switch (iters) {
      case 1: S3D_LN_PIPE(1); break;
      case 2: S3D_LN_PIPE(2); break;
      case 3: S3D_LN_PIPE(3); break;
      case 4: S3D_LN_PIPE(4); break;
      case 5: S3D_LN_PIPE(5); break;
      case 6: S3D_LN_PIPE(6); break;
      case 7: S3D_LN_PIPE(7); break;
      default: S3D_LN_PIPE(8); break;
    }
#undef S3D_LN_PIPE
    if (rc != S3D_ERR_UNSUPPORTED) return rc;
  }
#define S3D_LN_BWD(I)                                                                                                 \
  do {                                                                                                                \
    if (dy_is_bf16)                                                                                                   \
      S3D_CUDA_OK(launch_pdl(layernorm_bwd_kernel<I, true>, dim3((int)blocks), dim3(warps_per_block * 32), (size_t)(shmem), stream, dy, x, gamma, mean, rstd,  \
                                                                                          dres, dx, dxb, dgamma,      \
                                                                                          dbeta, dxsum, T, D));        \
    else                                                                                                              \
      S3D_CUDA_OK(launch_pdl(layernorm_bwd_kernel<I, false>, dim3((int)blocks), dim3(warps_per_block * 32), (size_t)(shmem), stream, dy, x, gamma, mean, rstd, \
                                                                                           dres, dx, dxb, dgamma,     \
                                                                                           dbeta, dxsum, T, D));       \
  } while (0)
  switch (iters) {
    case 1: S3D_LN_BWD(1); break;
    case 2: S3D_LN_BWD(2); break;
    case 3: S3D_LN_BWD(3); break;
    case 4: S3D_LN_BWD(4); break;
    case 5: S3D_LN_BWD(5); break;
    case 6: S3D_LN_BWD(6); break;
    case 7: S3D_LN_BWD(7); break;
    default: S3D_LN_BWD(8); break;
  }
#undef S3D_LN_BWD
  S3D_LAUNCH_OK();
  return S3D_OK;
}

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 cast (weight shadow copies / activations), with optional transposed copy [C,R] of a [R,C] matrix.
// ------------------------------------------------------------------------------------------------
__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n4) {
  pdl_prologue();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(in)[i];
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(out)[i] = u;
  }
}
__global__ void cast_bf16_tail_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t start,
                                      size_t n) {
  pdl_prologue();
  const size_t i = start + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16(in[i]);
}

int cast_f32_to_bf16(const float* in, void* out, long long n, cudaStream_t stream) {
  if (n <= 0) return S3D_ERR_BAD_SHAPE;
  if (in == nullptr || out == nullptr) return S3D_ERR_NULL;
  auto o = reinterpret_cast<__nv_bfloat16*>(out);
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
  const size_t n4 = aligned ? (size_t)n / 4 : 0;
  if (n4 > 0) {
    size_t blocks = (n4 + 255) / 256;
    const size_t cap = (size_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    S3D_CUDA_OK(launch_pdl(cast_bf16_kernel, dim3((int)blocks), dim3(256), (size_t)(0), stream, in, o, n4));
    S3D_LAUNCH_OK();
  }
  const size_t done = n4 * 4;
  if (done < (size_t)n) {
    const size_t rem = (size_t)n - done;
    S3D_CUDA_OK(launch_pdl(cast_bf16_tail_kernel, dim3((int)((rem + 255) / 256)), dim3(256), (size_t)(0), stream, in, o, done, (size_t)n));
    S3D_LAUNCH_OK();
  }
  return S3D_OK;
}

// out[c, r] = bf16(in[r, c]);  32x32 smem tiles, coalesced on both sides.
template <typename TIn>
__global__ void transpose_to_bf16_kernel(const TIn* __restrict__ in, __nv_bfloat16* __restrict__ out, int R, int C,
                                         long long ld_in, long long ld_out) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[j][threadIdx.x] = (float)in[(size_t)r * ld_in + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < C) out[(size_t)c * ld_out + r] = __float2bfloat16(tile[threadIdx.x][j]);
  }
}

int transpose_to_bf16(const void* in, int in_is_bf16, void* out, int R, int C, long long ld_in, long long ld_out,
                      cudaStream_t stream) {
  if (R <= 0 || C <= 0) return S3D_ERR_BAD_SHAPE;
  if (in == nullptr || out == nullptr) return S3D_ERR_NULL;
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  if (grid.y > 65535) return S3D_ERR_BAD_SHAPE;
  if (in_is_bf16)
    S3D_CUDA_OK(launch_pdl(transpose_to_bf16_kernel<__nv_bfloat16>, dim3(grid), dim3(block), (size_t)(0), stream, reinterpret_cast<const __nv_bfloat16*>(in),
                                                                        reinterpret_cast<__nv_bfloat16*>(out), R, C,
                                                                        ld_in, ld_out));
  else
    S3D_CUDA_OK(launch_pdl(transpose_to_bf16_kernel<float>, dim3(grid), dim3(block), (size_t)(0), stream, reinterpret_cast<const float*>(in),
                                                                reinterpret_cast<__nv_bfloat16*>(out), R, C, ld_in,
                                                                ld_out));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

// ------------------------------------------------------------------------------------------------
// Column sums of a bf16 [T, C] matrix into fp32 [C] (Linear bias gradients). out must be zeroed by the caller
// when accumulate == 0 is requested we zero it here. Each CTA covers 64 columns x a row slab; atomics per CTA.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out,
                                                         int T, int C, long long ld, int rows_per_block) {
  pdl_prologue();
  // blockDim = (32 column-pairs, 8 row lanes)
  __shared__ float red[8][64];
  const int c = blockIdx.x * 64 + threadIdx.x * 2;
  const int r_begin = blockIdx.y * rows_per_block;
  const int r_end = min(T, r_begin + rows_per_block);
  float s0 = 0.f, s1 = 0.f;
  if (c < C) {
    for (int r = r_begin + threadIdx.y; r < r_end; r += 8) {
      const uint32_t u = *reinterpret_cast<const uint32_t*>(in + (size_t)r * ld + c);
      const float2 f = unpack_bf16x2(u);
      s0 += f.x;
      s1 += f.y;
    }
  }
  red[threadIdx.y][threadIdx.x * 2] = s0;
  red[threadIdx.y][threadIdx.x * 2 + 1] = s1;
  __syncthreads();
  if (threadIdx.y == 0) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a += red[j][threadIdx.x * 2];
      b += red[j][threadIdx.x * 2 + 1];
    }
    if (c < C) atomicAdd(out + c, a);
    if (c + 1 < C) atomicAdd(out + c + 1, b);
  }
}

// 16-byte variant (C % 8 == 0, 16-byte aligned rows): a warp reads 512 contiguous bytes (256 columns) of a row, the 8
// warps of the CTA take rows r, r+8, ... four at a time (4 independent 16-byte loads per thread in flight).
__global__ void __launch_bounds__(256) colsum_bf16_v8_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out,
                                                            int T, int C, long long ld, int rows_per_block) {
  pdl_prologue();
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  const int r_begin = blockIdx.y * rows_per_block;
  const int r_end = min(T, r_begin + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    int r = r_begin + w;
    for (; r + 24 < r_end; r += 32) {
      uint4 u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) u[j] = *reinterpret_cast<const uint4*>(in + (size_t)(r + 8 * j) * ld + c);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf16x2(u[j].x), b = unpack_bf16x2(u[j].y), d = unpack_bf16x2(u[j].z),
                     e = unpack_bf16x2(u[j].w);
        acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
        acc[4] += d.x; acc[5] += d.y; acc[6] += e.x; acc[7] += e.y;
      }
    }
    for (; r < r_end; r += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(in + (size_t)r * ld + c);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), d = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
      acc[4] += d.x; acc[5] += d.y; acc[6] += e.x; acc[7] += e.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[w][lane * 8 + j] = acc[j];
  __syncthreads();
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col < C) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    atomicAdd(out + col, t);
  }
}

int colsum_bf16(const void* in, float* out, int T, int C, long long ld, int accumulate, cudaStream_t stream) {
  if (T <= 0 || C <= 0 || C % 2 != 0 || ld % 2 != 0) return S3D_ERR_BAD_SHAPE;
  if (in == nullptr || out == nullptr) return S3D_ERR_NULL;
  if (!accumulate) S3D_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)C, stream));
  if (C % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && T >= 256) {
    const int col_blocks = (C + 255) / 256;
    int row_blocks = (num_sms() * 8 + col_blocks - 1) / col_blocks;
    if (row_blocks > (T + 63) / 64) row_blocks = (T + 63) / 64;
    if (row_blocks < 1) row_blocks = 1;
    const int rows_per_block = (T + row_blocks - 1) / row_blocks;
    dim3 grid(col_blocks, row_blocks);
    S3D_CUDA_OK(launch_pdl(colsum_bf16_v8_kernel, dim3(grid), dim3(256), (size_t)(0), stream, reinterpret_cast<const __nv_bfloat16*>(in), out, T, C, ld,
                                                    rows_per_block));
    S3D_LAUNCH_OK();
    return S3D_OK;
  }
  const int col_blocks = (C + 63) / 64;
  int row_blocks = (num_sms() * 4 + col_blocks - 1) / col_blocks;
  if (row_blocks > (T + 63) / 64) row_blocks = (T + 63) / 64;
  if (row_blocks < 1) row_blocks = 1;
  const int rows_per_block = (T + row_blocks - 1) / row_blocks;
  dim3 grid(col_blocks, row_blocks), block(32, 8);
  S3D_CUDA_OK(launch_pdl(colsum_bf16_kernel, dim3(grid), dim3(block), (size_t)(0), stream, reinterpret_cast<const __nv_bfloat16*>(in), out, T, C, ld,
                                                 rows_per_block));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

// ------------------------------------------------------------------------------------------------
// Element-wise dropout of the group_embed layer (nn.TransformerEncoderLayer dropout1 / dropout / dropout2):
//   dropout_add : out = residual + keep(row, col) * x / (1 - p)      (fp32 residual stream)
//   dropout_bf16: out = keep(row, col) * x / (1 - p)                 (bf16 GEMM operand; in place allowed)
// The mask is the counter-based hash of common.cuh keyed by a device-resident seed (CUDA-graph replays see a new seed).
// ------------------------------------------------------------------------------------------------
__global__ void dropout_add_f32_kernel(const float* __restrict__ x, const float* __restrict__ res, float* __restrict__ out,
                                       long long n4, int cols, const uint32_t* __restrict__ seed, uint32_t site,
                                       uint32_t thresh14, float scale) {
  pdl_prologue();
  const uint32_t ss = drop_site_seed(*seed, site);
  const int c4 = cols >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t row = (uint32_t)(i / c4), col = (uint32_t)(i % c4) * 4u;
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    float4 o = res != nullptr ? reinterpret_cast<const float4*>(res)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t y = drop_rowkey(ss, row) + (col >> 1) * kDropColMul;
    const uint32_t h0 = drop_word(y), h1 = drop_word(y + kDropColMul);
    if ((h0 & 0xffffu) >= thresh14) o.x += v.x * scale;
    if ((h0 >> 16) >= thresh14) o.y += v.y * scale;
    if ((h1 & 0xffffu) >= thresh14) o.z += v.z * scale;
    if ((h1 >> 16) >= thresh14) o.w += v.w * scale;
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

__global__ void dropout_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, long long n4,
                                    int cols, const uint32_t* __restrict__ seed, uint32_t site, uint32_t thresh14,
                                    float scale) {
  pdl_prologue();
  const uint32_t ss = drop_site_seed(*seed, site);
  const int c4 = cols >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t row = (uint32_t)(i / c4), col = (uint32_t)(i % c4) * 4u;
    const uint2 u = reinterpret_cast<const uint2*>(x)[i];
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    const uint32_t y = drop_rowkey(ss, row) + (col >> 1) * kDropColMul;
    const uint32_t h0 = drop_word(y), h1 = drop_word(y + kDropColMul);
    uint2 w;
    w.x = pack_bf16x2((h0 & 0xffffu) >= thresh14 ? a.x * scale : 0.f, (h0 >> 16) >= thresh14 ? a.y * scale : 0.f);
    w.y = pack_bf16x2((h1 & 0xffffu) >= thresh14 ? b.x * scale : 0.f, (h1 >> 16) >= thresh14 ? b.y * scale : 0.f);
    reinterpret_cast<uint2*>(out)[i] = w;
  }
}

int dropout_add_f32(const float* x, const float* res, float* out, long long rows, int cols, const uint32_t* seed,
                    unsigned site, float p, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0 || cols % 4 != 0 || p < 0.f || p >= 1.f) return S3D_ERR_BAD_SHAPE;
  if (x == nullptr || out == nullptr || seed == nullptr) return S3D_ERR_NULL;
  const uint32_t th = drop_thresh14(p);
  const long long n4 = rows * cols / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  S3D_CUDA_OK(launch_pdl(dropout_add_f32_kernel, dim3((int)blocks), dim3(256), (size_t)(0), stream, x, res, out, n4, cols, seed, site, th,
                                                          drop_keep_scale(th)));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int dropout_bf16(const void* x, void* out, long long rows, int cols, const uint32_t* seed, unsigned site, float p,
                 cudaStream_t stream) {
  if (rows <= 0 || cols <= 0 || cols % 4 != 0 || p < 0.f || p >= 1.f) return S3D_ERR_BAD_SHAPE;
  if (x == nullptr || out == nullptr || seed == nullptr) return S3D_ERR_NULL;
  const uint32_t th = drop_thresh14(p);
  const long long n4 = rows * cols / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  S3D_CUDA_OK(launch_pdl(dropout_bf16_kernel, dim3((int)blocks), dim3(256), (size_t)(0), stream, reinterpret_cast<const __nv_bfloat16*>(x),
                                                       reinterpret_cast<__nv_bfloat16*>(out), n4, cols, seed, site, th,
                                                       drop_keep_scale(th)));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

// ------------------------------------------------------------------------------------------------
// Voxel patch gather: x [B,1,V,V,V] fp32 -> P bf16 [B*p*p*(zsum?1:p), Kpad], K = c^3 (zero padded to Kpad).
// Row order is (b, px, py, pz), column order (dx, dy, dz) = Conv3d weight [D,1,c,c,c] flattened, so the patchify
// conv (embed_layer_3d_modality.py:22-24) is P @ W^T. With zsum=1 the pz patches of a column are summed first
// (VoxelEmbed's mean over dim 4, :38, commutes with the linear map; the 1/p factor is applied in the GEMM epilogue).
// Occupancy values are 0/1 so sums <= p are exact in bf16.
// ------------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void __launch_bounds__(256) voxel_patch_gather_kernel(const TIn* __restrict__ x,
                                                                __nv_bfloat16* __restrict__ P, int B, int V, int c,
                                                                int p, int Kpad, int zsum) {
  pdl_prologue();
  const int K = c * c * c;
  const int pz_out = zsum ? 1 : p;
  const long long rows = (long long)B * p * p * pz_out;
  const int kp2 = Kpad >> 1;
  const long long total = rows * kp2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / kp2;
    const int k = (int)(i % kp2) * 2;
    long long t = row;
    const int pz = zsum ? 0 : (int)(t % p);
    if (!zsum) t /= p;
    const int py = (int)(t % p);
    t /= p;
    const int px = (int)(t % p);
    const int b = (int)(t / p);
    float v[2] = {0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int kk = k + e;
      if (kk < K) {
        const int dz = kk % c;
        const int dy = (kk / c) % c;
        const int dx = kk / (c * c);
        const size_t base = (((size_t)b * V + (size_t)(px * c + dx)) * V + (size_t)(py * c + dy)) * V;
        if (zsum) {
          float s = 0.f;
          for (int z = 0; z < p; ++z) s += (float)x[base + (size_t)(z * c + dz)];
          v[e] = s;
        } else {
          v[e] = (float)x[base + (size_t)(pz * c + dz)];
        }
      }
    }
    *reinterpret_cast<uint32_t*>(P + row * Kpad + k) = pack_bf16x2(v[0], v[1]);
  }
}

// Staged variant: one CTA per (sample, px, py) column of cells. The c * c z-lines of the column (V contiguous voxels each,
// read as whole lines) are parked in shared memory; the p output rows of the column (cells px, py, 0..p-1: p * Kpad
// CONTIGUOUS bf16) are then written with coalesced 4-byte stores, thread t producing the columns (2t, 2t + 1) of every
// row from a (line, dz) pair computed once. The element-wise kernel above spends ~10 integer divisions per output pair
// and reads single bytes (0.9 ms on the 64 x 128^3 batch against 0.06 ms of HBM time).
template <typename TIn>
__global__ void __launch_bounds__(256) voxel_patch_gather_staged_kernel(const TIn* __restrict__ x,
                                                                       __nv_bfloat16* __restrict__ P, int V, int c, int p,
                                                                       int Kpad, int zsum) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t vg_smem[];
  TIn* lines = reinterpret_cast<TIn*>(vg_smem);  // [c * c][V]
  const int K = c * c * c;
  const int py = blockIdx.x % p, px = (blockIdx.x / p) % p, b = blockIdx.x / (p * p);
  const int nlines = c * c;
  if ((V * sizeof(TIn)) % 16 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {  // whole lines as 16-byte vectors
    const int nvec = (int)(V * sizeof(TIn) / 16);
    for (int i = threadIdx.x; i < nlines * nvec; i += blockDim.x) {
      const int ln = i / nvec, q = i - ln * nvec;
      const int dx = ln / c, dy = ln - dx * c;
      const TIn* src = x + (((size_t)b * V + (size_t)(px * c + dx)) * V + (size_t)(py * c + dy)) * V;
      reinterpret_cast<uint4*>(lines + (size_t)ln * V)[q] = __ldg(reinterpret_cast<const uint4*>(src) + q);
    }
  } else {
    for (int i = threadIdx.x; i < nlines * V; i += blockDim.x) {
      const int ln = i / V, z = i - ln * V;
      const int dx = ln / c, dy = ln - dx * c;
      lines[i] = x[(((size_t)b * V + (size_t)(px * c + dx)) * V + (size_t)(py * c + dy)) * V + z];
    }
  }
  __syncthreads();
  const int kp2 = Kpad >> 1;
  for (int t = threadIdx.x; t < kp2; t += blockDim.x) {
    int off[2];  // line * V + dz of the two columns (-1: padding column)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int kk = 2 * t + e;
      const int ln = kk / c;
      off[e] = kk < K ? ln * V + (kk - ln * c) : -1;
    }
    if (zsum) {
      float v[2] = {0.f, 0.f};
#pragma unroll
      for (int e = 0; e < 2; ++e)
        if (off[e] >= 0)
          for (int z = 0; z < p; ++z) v[e] += (float)lines[off[e] + z * c];
      *reinterpret_cast<uint32_t*>(P + (size_t)blockIdx.x * Kpad + 2 * t) = pack_bf16x2(v[0], v[1]);
    } else {
      for (int pz = 0; pz < p; ++pz) {
        const float v0 = off[0] >= 0 ? (float)lines[off[0] + pz * c] : 0.f;
        const float v1 = off[1] >= 0 ? (float)lines[off[1] + pz * c] : 0.f;
        *reinterpret_cast<uint32_t*>(P + ((size_t)blockIdx.x * p + pz) * Kpad + 2 * t) = pack_bf16x2(v0, v1);
      }
    }
  }
}

template <typename TIn>
static int launch_voxel_gather_staged(const void* x, __nv_bfloat16* out, int B, int V, int cell, int patch, int Kpad,
                                      int zsum, cudaStream_t stream, bool* done) {
  const size_t shmem = (size_t)cell * cell * V * sizeof(TIn);
  const long long ctas = (long long)B * patch * patch;
  *done = false;
  if (shmem > 160 * 1024 || ctas > 0x7fffffffLL) return S3D_OK;  // the element-wise kernel handles it
  auto kern = voxel_patch_gather_staged_kernel<TIn>;
  S3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
  S3D_CUDA_OK(launch_pdl(kern, dim3((unsigned)ctas), dim3(256), shmem, stream, reinterpret_cast<const TIn*>(x), out, V, cell,
                         patch, Kpad, zsum));
  S3D_LAUNCH_OK();
  *done = true;
  return S3D_OK;
}

// in_dtype: 0 = float32, 1 = uint8 / bool occupancy, 2 = int32 (what the reference's binvox loaders yield,
// data/modelnet40.py:40, before `.float()` at train_cls_voxel.py:276)
int voxel_patch_gather(const void* x, int in_dtype, void* P, int B, int V, int cell, int patch, int Kpad, int zsum,
                       cudaStream_t stream) {
  if (B <= 0 || V <= 0 || cell <= 0 || patch <= 0 || patch * cell > V || Kpad < cell * cell * cell || Kpad % 8 != 0)
    return S3D_ERR_BAD_SHAPE;
  if (x == nullptr || P == nullptr) return S3D_ERR_NULL;
  if (in_dtype < 0 || in_dtype > 2) return S3D_ERR_UNSUPPORTED;
  const long long rows = (long long)B * patch * patch * (zsum ? 1 : patch);
  const long long total = rows * (Kpad / 2);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  auto out = reinterpret_cast<__nv_bfloat16*>(P);
  static const bool no_staged = getenv("S3D_VOXEL_GATHER_ELEMENTWISE") != nullptr;
  if (!no_staged) {
    bool done = false;
    int rc;
    if (in_dtype == 0) rc = launch_voxel_gather_staged<float>(x, out, B, V, cell, patch, Kpad, zsum, stream, &done);
    else if (in_dtype == 1) rc = launch_voxel_gather_staged<uint8_t>(x, out, B, V, cell, patch, Kpad, zsum, stream, &done);
    else rc = launch_voxel_gather_staged<int>(x, out, B, V, cell, patch, Kpad, zsum, stream, &done);
    if (rc) return rc;
    if (done) return S3D_OK;
  }
  if (in_dtype == 0)
    S3D_CUDA_OK(launch_pdl(voxel_patch_gather_kernel<float>, dim3((int)blocks), dim3(256), (size_t)(0), stream, reinterpret_cast<const float*>(x), out, B, V, cell, patch, Kpad, zsum));
  else if (in_dtype == 1)
    S3D_CUDA_OK(launch_pdl(voxel_patch_gather_kernel<uint8_t>, dim3((int)blocks), dim3(256), (size_t)(0), stream, reinterpret_cast<const uint8_t*>(x), out, B, V, cell, patch, Kpad, zsum));
  else
    S3D_CUDA_OK(launch_pdl(voxel_patch_gather_kernel<int>, dim3((int)blocks), dim3(256), (size_t)(0), stream, reinterpret_cast<const int*>(x), out, B, V, cell, patch, Kpad, zsum));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

// ------------------------------------------------------------------------------------------------
// Fused Adam (torch.optim.Adam semantics, no amsgrad, L2 weight decay folded into the gradient) over a flat fp32
// parameter segment; also refreshes the bf16 shadow copy used by the tensor-core GEMMs. grad_scale folds the
// 1/world_size of the data-parallel gradient average into the same pass.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                  float* __restrict__ m, float* __restrict__ v,
                                                  __nv_bfloat16* __restrict__ shadow, size_t n, float lr, float beta1,
                                                  float beta2, float eps, float weight_decay, float bias_corr1,
                                                  float bias_corr2_sqrt, float grad_scale,
                                                  const int* __restrict__ step_dev) {
  pdl_prologue();
  if (step_dev != nullptr) {  // CUDA-graph replays: the step counter lives on the device
    const float st = (float)(*step_dev);
    bias_corr1 = 1.f - powf(beta1, st);
    bias_corr2_sqrt = sqrtf(1.f - powf(beta2, st));
  }
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    const float pi = p[i];
    if (weight_decay != 0.f) gi += weight_decay * pi;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bias_corr2_sqrt + eps;
    const float pn = pi - (lr / bias_corr1) * (mi / denom);
    p[i] = pn;
    if (shadow != nullptr) shadow[i] = __float2bfloat16(pn);
  }
}

int adam_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, long long n, float lr, float beta1,
              float beta2, float eps, float weight_decay, int step, const int* step_dev, float grad_scale,
              cudaStream_t stream) {
  if (n <= 0 || (step <= 0 && step_dev == nullptr)) return S3D_ERR_BAD_SHAPE;
  if (p == nullptr || g == nullptr || m == nullptr || v == nullptr) return S3D_ERR_NULL;
  const float bc1 = 1.f - powf(beta1, (float)(step > 0 ? step : 1));
  const float bc2s = sqrtf(1.f - powf(beta2, (float)(step > 0 ? step : 1)));
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  S3D_CUDA_OK(launch_pdl(adam_kernel, dim3((int)blocks), dim3(256), (size_t)(0), stream, p, g, m, v, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), (size_t)n,
                                               lr, beta1, beta2, eps, weight_decay, bc1, bc2s, grad_scale, step_dev));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

// ------------------------------------------------------------------------------------------------
// Fused SGD with momentum (torch.optim.SGD semantics, dampening 0, no nesterov): the optimizer of the point scripts
// (train_cls.py:91, train_partseg.py:95). buf = mu * buf + g (buf = g on the first step); p -= lr * buf.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                 float* __restrict__ buf, __nv_bfloat16* __restrict__ shadow, size_t n,
                                                 float lr, float momentum, float weight_decay, float grad_scale,
                                                 int first_step, const int* __restrict__ step_dev) {
  pdl_prologue();
  if (step_dev != nullptr) first_step = (*step_dev <= 1);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float pi = p[i];
    float gi = g[i] * grad_scale;
    if (weight_decay != 0.f) gi += weight_decay * pi;
    const float b = first_step ? gi : momentum * buf[i] + gi;
    buf[i] = b;
    const float pn = pi - lr * b;
    p[i] = pn;
    if (shadow != nullptr) shadow[i] = __float2bfloat16(pn);
  }
}

int sgd_momentum_step(float* p, const float* g, float* buf, void* shadow_bf16, long long n, float lr, float momentum,
                      float weight_decay, int step, const int* step_dev, float grad_scale, cudaStream_t stream) {
  if (n <= 0 || (step <= 0 && step_dev == nullptr)) return S3D_ERR_BAD_SHAPE;
  if (p == nullptr || g == nullptr || buf == nullptr) return S3D_ERR_NULL;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  S3D_CUDA_OK(launch_pdl(sgd_kernel, dim3((int)blocks), dim3(256), (size_t)(0), stream, p, g, buf, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), (size_t)n, lr, momentum,
                                              weight_decay, grad_scale, step == 1, step_dev));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

}  // namespace s3d
