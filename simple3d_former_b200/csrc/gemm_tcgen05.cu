// Persistent, warp-specialised bf16 GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   D[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) )
//
// A and B are each either K-major (k contiguous in HBM: activations x weights of a Linear layer, reference
// timm-0.3.2 Attention.qkv / proj / Mlp.fc1 / fc2) or MN-major (m or n contiguous: used for the weight-gradient
// GEMMs dW = dY^T X where the contraction runs over tokens, so no transposed copies of activations are made).
//
// Roles (384 threads, one CTA per SM, grid = min(tiles, #SM)):
//   warp 0      : TMA producer (one elected lane), kStages-deep smem ring, 128B swizzle
//   warp 1      : MMA issuer (one elected lane), tcgen05.mma cta_group::1, M=128, N=BN, K=16, fp32 accum in TMEM
//   warp 2      : TMEM allocator (2 accumulator stages of BN columns)
//   warps 4..11 : epilogue; TMEM -> registers (tcgen05.ld 32x32b.x32) -> bias / GELU / dGELU / residual -> HBM
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "kernels.h"

#include <stdlib.h>

namespace s3d {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kNumThreads = 384;
constexpr int kEpiWarp0 = 4;

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, int A_MN, int B_MN>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int batch = blockIdx.y;

  const int num_m = (p.M + BM - 1) / BM;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 8);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------ TMA producer ------------------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * BM;
        const int n0 = (tile % num_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          const int k0 = kb * BK;
          if (p.batched) {
            if (A_MN) {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i) tma_load_3d(sa + i * 8192, &tma_a, &full_bar[stage], m0 + 64 * i, k0, batch);
            } else {
              tma_load_3d(sa, &tma_a, &full_bar[stage], k0, m0, batch);
            }
            if (B_MN) {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i) tma_load_3d(sb + i * 8192, &tma_b, &full_bar[stage], n0 + 64 * i, k0, batch);
            } else {
              tma_load_3d(sb, &tma_b, &full_bar[stage], k0, n0, batch);
            }
          } else {
            if (A_MN) {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i) tma_load_2d(sa + i * 8192, &tma_a, &full_bar[stage], m0 + 64 * i, k0);
            } else {
              tma_load_2d(sa, &tma_a, &full_bar[stage], k0, m0);
            }
            if (B_MN) {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i) tma_load_2d(sb + i * 8192, &tma_b, &full_bar[stage], n0 + 64 * i, k0);
            } else {
              tma_load_2d(sb, &tma_b, &full_bar[stage], k0, n0);
            }
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------- MMA issuer -------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = A_MN ? make_smem_desc_sw128(sa + k * 2048, p.mn_lbo, p.mn_sbo)
                                        : make_smem_desc_sw128(sa + k * 32, p.k_lbo, p.k_sbo);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128(sb + k * 2048, p.mn_lbo, p.mn_sbo)
                                        : make_smem_desc_sw128(sb + k * 32, p.k_lbo, p.k_sbo);
            umma_f16_ss(d_tmem, adesc, bdesc, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[as]);  // accumulator complete -> epilogue
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= kEpiWarp0) {
    // -------------------------------------------- epilogue --------------------------------------------
    const int e = warp - kEpiWarp0;  // 0..7
    const int quad = warp & 3;       // TMEM lane quadrant this warp may access
    const int half = e >> 2;         // which half of the BN columns
    constexpr int kColsPerHalf = BN / 2;
    int as = 0;
    uint32_t aphase = 0;
    const long long boff_d = p.batched ? (long long)batch * p.batch_stride_d : 0;
    const long long boff_r = p.batched ? (long long)batch * p.batch_stride_r : 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * BM;
      const int n0 = (tile % num_n) * BN;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < p.M;
#pragma unroll 1
      for (int c = 0; c < kColsPerHalf; c += 32) {
        const int col0 = half * kColsPerHalf + c;
        const int n = n0 + col0;
        if (n >= p.N) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + col0), v);
        tc_wait_ld();
        const bool full_chunk = (n + 32 <= p.N);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * p.alpha;
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full_chunk || n + j < p.N) f[j] += __ldg(p.bias + n + j);
        }
        if (p.epilogue == EPI_GELU) {
          if (p.aux_out != nullptr && row_ok) {
            __nv_bfloat16* ao = p.aux_out + (long long)row * p.ld_aux_out + n;
            if (full_chunk) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 u;
                u.x = pack_bf16x2(f[j], f[j + 1]);
                u.y = pack_bf16x2(f[j + 2], f[j + 3]);
                u.z = pack_bf16x2(f[j + 4], f[j + 5]);
                u.w = pack_bf16x2(f[j + 6], f[j + 7]);
                *reinterpret_cast<uint4*>(ao + j) = u;
              }
            } else {
              for (int j = 0; j < 32 && n + j < p.N; ++j) ao[j] = __float2bfloat16(f[j]);
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
        } else if (p.epilogue == EPI_DGELU) {
          if (row_ok) {
            const __nv_bfloat16* ai = p.aux_in + (long long)row * p.ld_aux_in + n;
            if (full_chunk) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint4 u = *reinterpret_cast<const uint4*>(ai + j);
                float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
                f[j] *= gelu_erf_grad(a0.x);
                f[j + 1] *= gelu_erf_grad(a0.y);
                f[j + 2] *= gelu_erf_grad(a1.x);
                f[j + 3] *= gelu_erf_grad(a1.y);
                f[j + 4] *= gelu_erf_grad(a2.x);
                f[j + 5] *= gelu_erf_grad(a2.y);
                f[j + 6] *= gelu_erf_grad(a3.x);
                f[j + 7] *= gelu_erf_grad(a3.y);
              }
            } else {
              for (int j = 0; j < 32 && n + j < p.N; ++j) f[j] *= gelu_erf_grad(__bfloat162float(ai[j]));
            }
          }
        }
        else if (p.epilogue == EPI_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        } else if (p.epilogue == EPI_DRELU) {
          // aux_in holds the post-ReLU activation: gradient passes where it is > 0
          if (row_ok) {
            const __nv_bfloat16* ai = p.aux_in + (long long)row * p.ld_aux_in + n;
            if (full_chunk) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint4 u = *reinterpret_cast<const uint4*>(ai + j);
                float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
                f[j] = a0.x > 0.f ? f[j] : 0.f;
                f[j + 1] = a0.y > 0.f ? f[j + 1] : 0.f;
                f[j + 2] = a1.x > 0.f ? f[j + 2] : 0.f;
                f[j + 3] = a1.y > 0.f ? f[j + 3] : 0.f;
                f[j + 4] = a2.x > 0.f ? f[j + 4] : 0.f;
                f[j + 5] = a2.y > 0.f ? f[j + 5] : 0.f;
                f[j + 6] = a3.x > 0.f ? f[j + 6] : 0.f;
                f[j + 7] = a3.y > 0.f ? f[j + 7] : 0.f;
              }
            } else {
              for (int j = 0; j < 32 && n + j < p.N; ++j) f[j] = __bfloat162float(ai[j]) > 0.f ? f[j] : 0.f;
            }
          }
        }
        if (row_ok) {
          if (p.residual != nullptr) {
            const float* r = p.residual + boff_r + (long long)row * p.ldr + n;
            if (full_chunk) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 rv = *reinterpret_cast<const float4*>(r + j);
                f[j] += rv.x;
                f[j + 1] += rv.y;
                f[j + 2] += rv.z;
                f[j + 3] += rv.w;
              }
            } else {
              for (int j = 0; j < 32 && n + j < p.N; ++j) f[j] += r[j];
            }
          }
          if (p.out_fp32) {
            float* d = reinterpret_cast<float*>(p.D) + boff_d + (long long)row * p.ldd + n;
            if (full_chunk) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(d + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
              for (int j = 0; j < 32 && n + j < p.N; ++j) d[j] = f[j];
            }
          } else {
            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.D) + boff_d + (long long)row * p.ldd + n;
            if (full_chunk) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 u;
                u.x = pack_bf16x2(f[j], f[j + 1]);
                u.y = pack_bf16x2(f[j + 2], f[j + 3]);
                u.z = pack_bf16x2(f[j + 4], f[j + 5]);
                u.w = pack_bf16x2(f[j + 6], f[j + 7]);
                *reinterpret_cast<uint4*>(d + j) = u;
              }
            } else {
              for (int j = 0; j < 32 && n + j < p.N; ++j) d[j] = __float2bfloat16(f[j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                      uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return S3D_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch_elems * 2) % 16 != 0) return S3D_ERR_ALIGNMENT;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? S3D_OK : S3D_ERR_DRIVER;
}

int make_tmap_bf16_3d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t batch,
                      uint64_t pitch_elems, uint64_t batch_pitch_elems, uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return S3D_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch_elems * 2) % 16 != 0 || (batch_pitch_elems * 2) % 16 != 0)
    return S3D_ERR_ALIGNMENT;
  cuuint64_t dims[3] = {inner, outer, batch};
  cuuint64_t strides[2] = {pitch_elems * 2, batch_pitch_elems * 2};
  cuuint32_t box[3] = {box_inner, box_outer, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? S3D_OK : S3D_ERR_DRIVER;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int BN, int A_MN, int B_MN>
static int launch_gemm(const GemmArgs& g, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap ta, tb;
  int rc;
  const GemmParams& p = g.p;
  if (g.batch > 1) {
    if (A_MN) rc = make_tmap_bf16_3d(&ta, g.A, p.M, p.K, g.batch, g.lda, g.batch_stride_a, 64, 64);
    else rc = make_tmap_bf16_3d(&ta, g.A, p.K, p.M, g.batch, g.lda, g.batch_stride_a, 64, BM);
    if (rc) return rc;
    if (B_MN) rc = make_tmap_bf16_3d(&tb, g.B, p.N, p.K, g.batch, g.ldb, g.batch_stride_b, 64, 64);
    else rc = make_tmap_bf16_3d(&tb, g.B, p.K, p.N, g.batch, g.ldb, g.batch_stride_b, 64, BN);
    if (rc) return rc;
  } else {
    if (A_MN) rc = make_tmap_bf16_2d(&ta, g.A, p.M, p.K, g.lda, 64, 64);
    else rc = make_tmap_bf16_2d(&ta, g.A, p.K, p.M, g.lda, 64, BM);
    if (rc) return rc;
    if (B_MN) rc = make_tmap_bf16_2d(&tb, g.B, p.N, p.K, g.ldb, 64, 64);
    else rc = make_tmap_bf16_2d(&tb, g.B, p.K, p.N, g.ldb, 64, BN);
    if (rc) return rc;
  }
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    S3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  int gx = tiles;
  const int sms = num_sms();
  const int per_batch_cap = g.batch > 1 ? (sms / g.batch > 0 ? sms / g.batch : 1) : sms;
  if (gx > per_batch_cap) gx = per_batch_cap;
  dim3 grid(gx, g.batch > 1 ? g.batch : 1);
  kern<<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  S3D_LAUNCH_OK();
  return S3D_OK;
}

template <int BN>
static int dispatch_major(const GemmArgs& g, cudaStream_t s) {
  if (g.a_mn == 0 && g.b_mn == 0) return launch_gemm<BN, 0, 0>(g, s);
  if (g.a_mn == 1 && g.b_mn == 1) return launch_gemm<BN, 1, 1>(g, s);
  if (g.a_mn == 0 && g.b_mn == 1) return launch_gemm<BN, 0, 1>(g, s);
  return launch_gemm<BN, 1, 0>(g, s);
}

static unsigned env_u32(const char* name, unsigned dflt) {
  const char* v = getenv(name);
  return v ? (unsigned)strtoul(v, nullptr, 0) : dflt;
}

int gemm_bf16(const GemmArgs& g_in, cudaStream_t stream) {
  static const unsigned mn_lbo = env_u32("S3D_DBG_MN_LBO", 8192), mn_sbo = env_u32("S3D_DBG_MN_SBO", 1024);
  static const unsigned k_lbo = env_u32("S3D_DBG_K_LBO", 16), k_sbo = env_u32("S3D_DBG_K_SBO", 1024);
  GemmArgs g = g_in;
  g.p.mn_lbo = mn_lbo;
  g.p.mn_sbo = mn_sbo;
  g.p.k_lbo = k_lbo;
  g.p.k_sbo = k_sbo;
  const GemmParams& p = g.p;
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return S3D_ERR_BAD_SHAPE;
  if (g.A == nullptr || g.B == nullptr || p.D == nullptr) return S3D_ERR_NULL;
  if ((p.epilogue == EPI_DGELU || p.epilogue == EPI_DRELU) && p.aux_in == nullptr) return S3D_ERR_NULL;
  // vector epilogue accesses need 16-byte aligned rows
  const long long dmul = p.out_fp32 ? 4 : 8;
  if (p.ldd % dmul != 0 || (reinterpret_cast<uintptr_t>(p.D) & 15) != 0) return S3D_ERR_ALIGNMENT;
  if (p.residual != nullptr && (p.ldr % 4 != 0 || (reinterpret_cast<uintptr_t>(p.residual) & 15) != 0))
    return S3D_ERR_ALIGNMENT;
  if (p.aux_in != nullptr && (p.ld_aux_in % 8 != 0 || (reinterpret_cast<uintptr_t>(p.aux_in) & 15) != 0))
    return S3D_ERR_ALIGNMENT;
  if (p.aux_out != nullptr && (p.ld_aux_out % 8 != 0 || (reinterpret_cast<uintptr_t>(p.aux_out) & 15) != 0))
    return S3D_ERR_ALIGNMENT;
  int bn = g.force_bn;
  if (bn == 0) {
    // BN=256 halves the smem operand traffic per flop; use it only when it still fills the machine.
    const long long t256 = (long long)((p.M + BM - 1) / BM) * ((p.N + 255) / 256) * (g.batch > 1 ? g.batch : 1);
    const long long t128 = (long long)((p.M + BM - 1) / BM) * ((p.N + 127) / 128) * (g.batch > 1 ? g.batch : 1);
    if (p.N > 128 && t256 >= num_sms()) bn = 256;
    else if (p.N > 64 && t128 >= num_sms() / 2) bn = 128;
    else if (p.N > 64 && p.N % 128 == 0) bn = 128;
    else bn = (p.N > 64 && t128 * 2 > num_sms()) ? 128 : 64;
  }
  switch (bn) {
    case 256: return dispatch_major<256>(g, stream);
    case 128: return dispatch_major<128>(g, stream);
    case 64: return dispatch_major<64>(g, stream);
    default: return S3D_ERR_UNSUPPORTED;
  }
}

}  // namespace s3d
