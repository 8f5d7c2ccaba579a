// Flash attention forward on the 5th-gen tensor cores (tcgen05 + TMEM + TMA) for long sequences -- the group_embed
// layer of the reference (nn.TransformerEncoderLayer over S = B*196 = 12544 tokens, dh = 192, vit_3d_2d_pretrain.py:381,479),
// which the reference evaluates by materialising 15*4 score matrices of S x S fp32 (37.8 GB).
//
// One CTA owns TWO 128-row query tiles (A, B) of one (batch, head) and streams 64-key K/V tiles through a 2-stage TMA
// ring. Per tile and K/V block:   S = Q K^T  (tcgen05.mma M128 N64 K16 x DH/16, accumulator in TMEM)
//                                 P = exp2(S*c - m)  (softmax warps: tcgen05.ld -> registers -> bf16 -> swizzled smem)
//                                 O += P V   (tcgen05.mma M128 N=DH K16 x 4, V read as an MN-major B operand, O in TMEM)
// The two tiles ping-pong: while the softmax warps of tile A work on S_A, the tensor core runs S_B / P_B V, so the MMA
// pipe stays busy. Running maxima are updated lazily (rescale O only when the max grew by more than 2^8), which keeps the
// TMEM read-modify-write of O off the critical path; row sums absorb the rest exactly.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = softmax of tile A, warps 6-9 = tile B.
// TMEM (512 columns): O_A [0,DH)  O_B [DH,2DH)  S_A [2DH, 2DH+64)  S_B [2DH+64, 2DH+128).
#include "kernels.h"

namespace s3d {

constexpr int kFaBM = 128;   // query rows per tile
constexpr int kFaBN = 64;    // keys per K/V block
constexpr int kFaThreads = 320;
constexpr float kFaLog2e = 1.4426950408889634f;

template <int DH>
struct FaCfg {
  static constexpr int kQBytes = kFaBM * DH * 2;          // per query tile
  static constexpr int kKVBytes = kFaBN * DH * 2;         // K or V block
  static constexpr int kPBytes = kFaBM * kFaBN * 2;       // P tile (bf16)
  static constexpr int kSmemBytes = 2 * kQBytes + 4 * kKVBytes + 2 * kPBytes + 1024 + 256;
  static constexpr int kTmemCols = 512;
  static constexpr int kColS = 2 * DH;                    // first S column
};

struct FaParams {
  __nv_bfloat16* out;
  float* lse;
  int N, H;
  long long row_bs;   // rows of the 2-D qkv view per batch index (timm layout: N, sequence-first: 0)
  long long col_bs;   // columns per batch index (timm: 0, sequence-first: 3E)
  int col_q, col_k, col_v;  // column of head 0 for q / k / v
  long long o_bs, o_hs, o_rs;
  float scale;
};

template <int DH>
__global__ void __launch_bounds__(kFaThreads, 1)
fa_fwd_tc_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv, const FaParams p) {
  using Cfg = FaCfg<DH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // [2][DH/64][128][128B]
  uint8_t* sK = sQ + 2 * Cfg::kQBytes;                 // [2 stages][DH/64][64][128B]
  uint8_t* sV = sK + 2 * Cfg::kKVBytes;                // [2 stages][DH/64][64][128B]
  uint8_t* sP = sV + 2 * Cfg::kKVBytes;                // [2 tiles][128][128B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * Cfg::kPBytes);
  uint64_t* q_full = bars;          // [2]
  uint64_t* k_full = bars + 2;      // [2]
  uint64_t* v_full = bars + 4;      // [2]
  uint64_t* kv_empty = bars + 6;    // [2]
  uint64_t* s_full = bars + 8;      // [2]
  uint64_t* p_full = bars + 10;     // [2]
  uint64_t* o_full = bars + 12;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * 2 * kFaBM;  // first query row of tile A
  const int nkv = (p.N + kFaBN - 1) / kFaBN;
  const int row_base = (int)(b * p.row_bs);
  const int cq = (int)(b * p.col_bs) + p.col_q + h * DH;
  const int ck = (int)(b * p.col_bs) + p.col_k + h * DH;
  const int cv = (int)(b * p.col_bs) + p.col_v + h * DH;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // --------------------------------------------- TMA producer ---------------------------------------------
    if (lane == 0) {
      for (int t = 0; t < 2; ++t) {
        mbar_expect_tx(&q_full[t], Cfg::kQBytes);
#pragma unroll
        for (int c = 0; c < DH / 64; ++c)
          tma_load_2d(sQ + t * Cfg::kQBytes + c * (kFaBM * 128), &tma_q, &q_full[t], cq + 64 * c, row_base + q0 + t * kFaBM);
      }
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], Cfg::kKVBytes);
#pragma unroll
        for (int c = 0; c < DH / 64; ++c)
          tma_load_2d(sK + st * Cfg::kKVBytes + c * (kFaBN * 128), &tma_kv, &k_full[st], ck + 64 * c, row_base + j * kFaBN);
        mbar_expect_tx(&v_full[st], Cfg::kKVBytes);
#pragma unroll
        for (int c = 0; c < DH / 64; ++c)
          tma_load_2d(sV + st * Cfg::kKVBytes + c * (kFaBN * 128), &tma_kv, &v_full[st], cv + 64 * c, row_base + j * kFaBN);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------- MMA issuer ----------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kFaBM, kFaBN, 0, 0);  // S = Q K^T : both K-major
      constexpr uint32_t idesc_o = make_idesc_bf16(kFaBM, DH, 0, 1);     // O = P V   : V is MN-major
      auto issue_s = [&](int t, int st) {
        const uint32_t a0 = smem_u32(sQ + t * Cfg::kQBytes);
        const uint32_t b0 = smem_u32(sK + st * Cfg::kKVBytes);
        const uint32_t d = tmem_base + Cfg::kColS + t * kFaBN;
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) {
          const uint64_t ad = make_smem_desc_sw128(a0 + (kk >> 2) * (kFaBM * 128) + (kk & 3) * 32, 16, 1024);
          const uint64_t bd = make_smem_desc_sw128(b0 + (kk >> 2) * (kFaBN * 128) + (kk & 3) * 32, 16, 1024);
          umma_f16_ss(d, ad, bd, idesc_s, kk != 0);
        }
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int st, int j) {
        const uint32_t a0 = smem_u32(sP + t * Cfg::kPBytes);
        const uint32_t b0 = smem_u32(sV + st * Cfg::kKVBytes);
        const uint32_t d = tmem_base + t * DH;
#pragma unroll
        for (int kk = 0; kk < kFaBN / 16; ++kk) {
          const uint64_t ad = make_smem_desc_sw128(a0 + kk * 32, 16, 1024);
          const uint64_t bd = make_smem_desc_sw128(b0 + kk * 2048, kFaBN * 128, 1024);
          umma_f16_ss(d, ad, bd, idesc_o, (j > 0) || (kk != 0));
        }
      };
      mbar_wait(&k_full[0], 0);
      mbar_wait(&q_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      mbar_wait(&q_full[1], 0);
      tc_fence_after();
      issue_s(1, 0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        const uint32_t ph = j & 1;
        mbar_wait(&v_full[st], (j >> 1) & 1);
        if (j + 1 < nkv) mbar_wait(&k_full[st ^ 1], ((j + 1) >> 1) & 1);
        // tile A
        mbar_wait(&p_full[0], ph);
        tc_fence_after();
        issue_pv(0, st, j);
        if (j + 1 < nkv) issue_s(0, st ^ 1);
        else umma_commit(&o_full[0]);
        // tile B
        mbar_wait(&p_full[1], ph);
        tc_fence_after();
        issue_pv(1, st, j);
        umma_commit(&kv_empty[st]);  // every MMA that reads stage `st` has been issued; frees it when they retire
        if (j + 1 < nkv) issue_s(1, st ^ 1);
        else umma_commit(&o_full[1]);
      }
    }
    __syncwarp();
  } else {
    // ----------------------------------------------- softmax -----------------------------------------------
    const int t = (warp - 2) >> 2;     // tile 0 (A) / 1 (B)
    const int quad = warp & 3;         // TMEM lane quadrant accessible to this warp
    const int r = quad * 32 + lane;    // row within the tile
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + Cfg::kColS + t * kFaBN;
    const uint32_t o_addr = tmem_base + lane_addr + t * DH;
    uint8_t* prow = sP + t * Cfg::kPBytes + r * 128;
    const float c = p.scale * kFaLog2e;
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32b_x32(s_addr, v0);
      tmem_ld_32x32b_x32(s_addr + 32, v1);
      tc_wait_ld();
      float s[64];
#pragma unroll
      for (int i = 0; i < 32; ++i) { s[i] = __uint_as_float(v0[i]); s[32 + i] = __uint_as_float(v1[i]); }
      const int key0 = j * kFaBN;
      if (key0 + kFaBN > p.N) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (key0 + i >= p.N) s[i] = -INFINITY;
      }
      float mx = s[0];
#pragma unroll
      for (int i = 1; i < 64; ++i) mx = fmaxf(mx, s[i]);
      // lazy rescale: keep the reference max unless it grew by more than 2^8 (P stays <= 256, exact in the row sums)
      const bool need = (mx - m_ref) * c > 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float f = need ? fast_exp2((m_ref - mx) * c) : 1.0f;  // m_ref = -inf -> f = 0 (first block: O is overwritten)
        if (j > 0) {
#pragma unroll 1
          for (int cc = 0; cc < DH; cc += 32) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(o_addr + cc, o);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st_32x32b_x32(o_addr + cc, o);
          }
          tc_wait_st();
        }
        if (need) { l *= f; m_ref = mx; }
      }
      const float mc = m_ref * c;
      float sum = 0.f;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          e[i] = fast_exp2(s[ch * 8 + i] * c - mc);
          sum += e[i];
        }
        uint4 u;
        u.x = pack_bf16x2(e[0], e[1]);
        u.y = pack_bf16x2(e[2], e[3]);
        u.z = pack_bf16x2(e[4], e[5]);
        u.w = pack_bf16x2(e[6], e[7]);
        *reinterpret_cast<uint4*>(prow + ((ch ^ (r & 7)) << 4)) = u;
      }
      l += sum;
      fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(&p_full[t]);
    }
    // epilogue: O / l -> bf16, lse
    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    const int row = q0 + t * kFaBM + r;
    const float inv = 1.0f / l;
    __nv_bfloat16* orow = p.out + (long long)b * p.o_bs + (long long)h * p.o_hs + (long long)row * p.o_rs;
#pragma unroll 1
    for (int cc = 0; cc < DH; cc += 32) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(o_addr + cc, o);
      tc_wait_ld();
      if (row < p.N) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + cc + i) = u;
        }
      }
    }
    if (p.lse != nullptr && row < p.N) p.lse[((long long)b * p.H + h) * p.N + row] = m_ref * p.scale + logf(l);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
// Supported when q/k/v are slices of one row-major 2-D buffer: either the timm layout [B, N, 3, H, dh] (batch selects
// rows) or the sequence-first layout [S, Nb, 3, H, dh] (batch selects columns). Returns S3D_ERR_UNSUPPORTED otherwise
// so the caller can use the generic mma.sync kernel.
template <int DH>
static int fa_fwd_launch(const AttnParams& a, cudaStream_t stream) {
  using Cfg = FaCfg<DH>;
  const long long E = (long long)a.H * DH;
  const __nv_bfloat16* base = a.q;
  const long long koff = a.k - a.q, voff = a.v - a.q;
  if (koff != E || voff != 2 * E || a.qkv_hs != DH) return S3D_ERR_UNSUPPORTED;
  FaParams p{};
  long long rows_total, width;
  if (a.qkv_rs == 3 * E && (a.qkv_bs == (long long)a.N * 3 * E || a.B == 1)) {  // timm: [B*N, 3E]
    rows_total = (long long)a.B * a.N;
    width = 3 * E;
    p.row_bs = a.N;
    p.col_bs = 0;
  } else if (a.qkv_bs == 3 * E && a.qkv_rs == (long long)a.B * 3 * E) {  // sequence-first: [S, Nb*3E]
    rows_total = a.N;
    width = (long long)a.B * 3 * E;
    p.row_bs = 0;
    p.col_bs = 3 * E;
  } else {
    return S3D_ERR_UNSUPPORTED;
  }
  CUtensorMap tq, tkv;
  int rc = make_tmap_bf16_2d(&tq, base, (uint64_t)width, (uint64_t)rows_total, (uint64_t)a.qkv_rs, 64, kFaBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tkv, base, (uint64_t)width, (uint64_t)rows_total, (uint64_t)a.qkv_rs, 64, kFaBN);
  if (rc) return rc;
  p.out = a.out;
  p.lse = a.lse;
  p.N = a.N;
  p.H = a.H;
  p.col_q = 0;
  p.col_k = (int)E;
  p.col_v = (int)(2 * E);
  p.o_bs = a.o_bs;
  p.o_hs = a.o_hs;
  p.o_rs = a.o_rs;
  p.scale = a.scale;
  if ((a.o_rs % 8) || (a.o_hs % 8) || (a.o_bs % 8) || (reinterpret_cast<uintptr_t>(a.out) & 15)) return S3D_ERR_ALIGNMENT;
  auto kern = fa_fwd_tc_kernel<DH>;
  static bool attr_set = false;
  if (!attr_set) {
    S3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  if (a.B > 65535 || a.H > 65535) return S3D_ERR_BAD_SHAPE;
  dim3 grid((a.N + 2 * kFaBM - 1) / (2 * kFaBM), a.H, a.B);
  kern<<<grid, kFaThreads, Cfg::kSmemBytes, stream>>>(tq, tkv, p);
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int attn_fwd_tc(const AttnParams& p, int DH, cudaStream_t stream) {
  if (p.B <= 0 || p.H <= 0 || p.N <= 0) return S3D_ERR_BAD_SHAPE;
  switch (DH) {
    case 192: return fa_fwd_launch<192>(p, stream);
    case 64: return fa_fwd_launch<64>(p, stream);
    default: return S3D_ERR_UNSUPPORTED;
  }
}

}  // namespace s3d
