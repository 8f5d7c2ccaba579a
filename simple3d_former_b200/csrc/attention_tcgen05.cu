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
  const uint32_t* drop_seed;  // attention-probability dropout: device seed (nullptr = off)
  uint32_t drop_site, drop_thresh16;
  float drop_scale;           // 1 / (1 - p)
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
      mbar_init(&p_full[i], 4);  // one arrive per softmax warp
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
    // The whole warp runs the control flow (waits are warp-uniform); one elected lane issues. Descriptor low words are
    // precomputed so each tcgen05.mma costs a couple of integer adds (at N = 64 an MMA lasts only ~40 cycles, so the
    // issue path must stay far below that).
    {
      constexpr uint32_t idesc_s = make_idesc_bf16(kFaBM, kFaBN, 0, 0);  // S = Q K^T : both K-major
      constexpr uint32_t idesc_o = make_idesc_bf16(kFaBM, DH, 0, 1);     // O = P V   : V is MN-major
      constexpr uint32_t hi = smem_desc_hi_sw128(1024);
      const uint32_t q_lo = smem_desc_lo(smem_u32(sQ), 16), k_lo = smem_desc_lo(smem_u32(sK), 16);
      const uint32_t p_lo = smem_desc_lo(smem_u32(sP), 16), v_lo = smem_desc_lo(smem_u32(sV), kFaBN * 128);
      auto issue_s = [&](int t, int st) {
        const uint32_t a = q_lo + t * (Cfg::kQBytes >> 4), b = k_lo + st * (Cfg::kKVBytes >> 4);
        const uint32_t d = tmem_base + Cfg::kColS + t * kFaBN;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk)
            umma_f16_ss2(d, a + (((kk >> 2) * (kFaBM * 128) + (kk & 3) * 32) >> 4), hi,
                         b + (((kk >> 2) * (kFaBN * 128) + (kk & 3) * 32) >> 4), hi, idesc_s, kk != 0);
          umma_commit(&s_full[t]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int t, int st, int j, uint64_t* done0, uint64_t* done1) {
        const uint32_t a = p_lo + t * (Cfg::kPBytes >> 4), b = v_lo + st * (Cfg::kKVBytes >> 4);
        const uint32_t d = tmem_base + t * DH;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kFaBN / 16; ++kk)
            umma_f16_ss2(d, a + ((kk * 32) >> 4), hi, b + ((kk * 2048) >> 4), hi, idesc_o, (j > 0) || (kk != 0));
          if (done0 != nullptr) umma_commit(done0);
          if (done1 != nullptr) umma_commit(done1);
        }
        __syncwarp();
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
        const bool last = (j + 1 == nkv);
        mbar_wait(&v_full[st], (j >> 1) & 1);
        if (!last) mbar_wait(&k_full[st ^ 1], ((j + 1) >> 1) & 1);
        // tile A
        mbar_wait(&p_full[0], ph);
        tc_fence_after();
        issue_pv(0, st, j, last ? &o_full[0] : nullptr, nullptr);
        if (!last) issue_s(0, st ^ 1);
        // tile B (its P V is the last reader of K/V stage `st`: the commit frees the stage when those MMAs retire)
        mbar_wait(&p_full[1], ph);
        tc_fence_after();
        issue_pv(1, st, j, &kv_empty[st], last ? &o_full[1] : nullptr);
        if (!last) issue_s(1, st ^ 1);
      }
    }
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
    // dropout on the attention probabilities (MultiheadAttention(dropout=p) inside nn.TransformerEncoderLayer): P V uses
    // P o mask, the softmax normaliser the full row sum; mask element = (row (b*H + h)*N + query, column key)
    const bool drop = p.drop_seed != nullptr;
    const uint32_t drop_row = drop ? (drop_site_seed(*p.drop_seed, p.drop_site) ^
                                      (((uint32_t)(b * p.H + h) * (uint32_t)p.N + (uint32_t)(q0 + t * kFaBM + r)) * kDropRowMul))
                                   : 0u;
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
      float mx4[4] = {s[0], s[1], s[2], s[3]};  // 4 independent chains instead of one 64-long dependent chain
#pragma unroll
      for (int i = 4; i < 64; i += 4) {
        mx4[0] = fmaxf(mx4[0], s[i]);
        mx4[1] = fmaxf(mx4[1], s[i + 1]);
        mx4[2] = fmaxf(mx4[2], s[i + 2]);
        mx4[3] = fmaxf(mx4[3], s[i + 3]);
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
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
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xs = fmaf(s[ch * 8 + i], c, -mc);
          e[i] = (i & 1) ? poly_exp2(xs) : fast_exp2(xs);  // half on MUFU, half on the FMA pipe
          sum4[i & 3] += e[i];
        }
        if (drop) {
          const uint32_t cp0 = (uint32_t)(key0 + ch * 8) >> 1;
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            const uint32_t hsh = drop_mix(drop_row ^ ((cp0 + (i >> 1)) * kDropColMul));
            if ((hsh & 0xffffu) < p.drop_thresh16) e[i] = 0.f;
            if ((hsh >> 16) < p.drop_thresh16) e[i + 1] = 0.f;
          }
        }
        uint4 u;
        u.x = pack_bf16x2(e[0], e[1]);
        u.y = pack_bf16x2(e[2], e[3]);
        u.z = pack_bf16x2(e[4], e[5]);
        u.w = pack_bf16x2(e[6], e[7]);
        *reinterpret_cast<uint4*>(prow + ((ch ^ (r & 7)) << 4)) = u;
      }
      l += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }
    // epilogue: O / l -> bf16, lse
    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    const int row = q0 + t * kFaBM + r;
    const float inv = (drop ? p.drop_scale : 1.0f) / l;
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
  p.drop_seed = a.drop_seed;
  p.drop_site = a.drop_site;
  p.drop_thresh16 = a.drop_thresh16;
  p.drop_scale = a.drop_scale;
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

// ================================================================================================
// Backward on tcgen05. Two kernels (no atomics, no dQ round trips through HBM):
//   dQ   : CTA = 128 query rows; per 64-key block  S = Q K^T, dP = dO V^T (TMEM, double buffered),
//          dS = P o (dP - delta) -> bf16 smem,  dQ += dS K  (K block re-read as an MN-major B operand)
//   dKdV : CTA = 128 key rows; per 64-query block S^T = K Q^T, dP^T = V dO^T,  P^T / dS^T -> bf16 smem,
//          dV += P^T dO,  dK += dS^T Q  (Q / dO blocks re-read as MN-major B operands)
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 element-wise: a TMEM lane (= row) is shared by
// two threads (warps w and w+4 reach the same lane quadrant), each handling 32 of the 64 columns of a block -- the
// element-wise phase (exp2, dS, bf16 stores) is what paces these kernels, so it gets 8 warps.
// ================================================================================================
struct FaBwdParams {
  __nv_bfloat16 *dq, *dk, *dv;  // outputs, qkv strides
  const float* lse;
  const float* delta;
  int N, H;
  long long row_bs, col_bs;      // qkv 2-D view
  int col_q, col_k, col_v;
  long long o_row_bs, o_col_bs;  // dout 2-D view
  long long qkv_bs, qkv_hs, qkv_rs;
  float scale;
  const uint32_t* drop_seed;  // attention-probability dropout (same mask as the forward kernel)
  uint32_t drop_site, drop_thresh16;
  float drop_scale;
};

constexpr int kFaBwdThreads = 320;

template <int DH>
struct FaBwdCfg {
  static constexpr int kTileBytes = 128 * DH * 2;   // 128-row operand tile
  static constexpr int kBlkBytes = 64 * DH * 2;     // 64-row streamed block
  static constexpr int kSBytes = 128 * 64 * 2;      // bf16 P / dS tile
  static constexpr int kSmemBytes = 2 * kTileBytes + 4 * kBlkBytes + 2 * kSBytes + 1024 + 1024;
};

// 32 bf16 (half of a 128-byte row) into the 128B-swizzled tile: logical 16-byte chunks half*4 .. half*4+3 of row r
__device__ __forceinline__ void store_half_row_bf16_sw128(uint8_t* row_base, int r, int half, const float (&e)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack_bf16x2(e[q * 8 + 0], e[q * 8 + 1]);
    u.y = pack_bf16x2(e[q * 8 + 2], e[q * 8 + 3]);
    u.z = pack_bf16x2(e[q * 8 + 4], e[q * 8 + 5]);
    u.w = pack_bf16x2(e[q * 8 + 6], e[q * 8 + 7]);
    *reinterpret_cast<uint4*>(row_base + (((half * 4 + q) ^ (r & 7)) << 4)) = u;
  }
}

// ------------------------------------------------ dQ ------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(kFaBwdThreads, 1)
fa_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tma_q128, const __grid_constant__ CUtensorMap tma_kv64,
                    const __grid_constant__ CUtensorMap tma_do128, const FaBwdParams p) {
  using Cfg = FaBwdCfg<DH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                          // [DH/64][128][128B]
  uint8_t* sdO = sQ + Cfg::kTileBytes;         // [DH/64][128][128B]
  uint8_t* sK = sdO + Cfg::kTileBytes;         // [2][DH/64][64][128B]
  uint8_t* sV = sK + 2 * Cfg::kBlkBytes;       // [2][DH/64][64][128B]
  uint8_t* sdS = sV + 2 * Cfg::kBlkBytes;      // [2][128][128B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 2 * Cfg::kSBytes);
  uint64_t* qdo_full = bars;       // [1]
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* v_full = bars + 3;     // [2]
  uint64_t* kv_empty = bars + 5;   // [2]
  uint64_t* sp_full = bars + 7;    // [2]
  uint64_t* ds_full = bars + 9;    // [2]
  uint64_t* dq_full = bars + 11;   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * 128;
  const int nkv = (p.N + 63) / 64;
  const int row_base = (int)(b * p.row_bs);
  const int cq = (int)(b * p.col_bs) + p.col_q + h * DH;
  const int ck = (int)(b * p.col_bs) + p.col_k + h * DH;
  const int cv = (int)(b * p.col_bs) + p.col_v + h * DH;
  const int o_row_base = (int)(b * p.o_row_bs);
  const int co = (int)(b * p.o_col_bs) + h * DH;
  constexpr int kColS = DH;  // TMEM: dQ [0,DH)  then per buffer u: S [DH + 128u, +64), dP [DH + 128u + 64, +64)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q128);
    tma_prefetch_desc(&tma_kv64);
    tma_prefetch_desc(&tma_do128);
    mbar_init(qdo_full, 1);
    mbar_init(dq_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&sp_full[i], 1);
      mbar_init(&ds_full[i], 8);  // one arrive per element-wise warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(qdo_full, 2 * Cfg::kTileBytes);
#pragma unroll
      for (int c = 0; c < DH / 64; ++c) {
        tma_load_2d(sQ + c * (128 * 128), &tma_q128, qdo_full, cq + 64 * c, row_base + q0);
        tma_load_2d(sdO + c * (128 * 128), &tma_do128, qdo_full, co + 64 * c, o_row_base + q0);
      }
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], Cfg::kBlkBytes);
#pragma unroll
        for (int c = 0; c < DH / 64; ++c)
          tma_load_2d(sK + st * Cfg::kBlkBytes + c * (64 * 128), &tma_kv64, &k_full[st], ck + 64 * c, row_base + j * 64);
        mbar_expect_tx(&v_full[st], Cfg::kBlkBytes);
#pragma unroll
        for (int c = 0; c < DH / 64; ++c)
          tma_load_2d(sV + st * Cfg::kBlkBytes + c * (64 * 128), &tma_kv64, &v_full[st], cv + 64 * c, row_base + j * 64);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // warp-converged control flow, one elected lane issues; precomputed descriptor low words (see the forward kernel)
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_q = make_idesc_bf16(128, DH, 0, 1);
    constexpr uint32_t hi = smem_desc_hi_sw128(1024);
    const uint32_t q_lo = smem_desc_lo(smem_u32(sQ), 16), do_lo = smem_desc_lo(smem_u32(sdO), 16);
    const uint32_t k_lo = smem_desc_lo(smem_u32(sK), 16), v_lo = smem_desc_lo(smem_u32(sV), 16);
    const uint32_t ds_lo = smem_desc_lo(smem_u32(sdS), 16), kmn_lo = smem_desc_lo(smem_u32(sK), 64 * 128);
    auto issue_s_dp = [&](int u, int st) {
      const uint32_t bk = k_lo + st * (Cfg::kBlkBytes >> 4), bv = v_lo + st * (Cfg::kBlkBytes >> 4);
      const uint32_t ds_ = tmem_base + kColS + u * 128, dp_ = ds_ + 64;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_f16_ss2(ds_, q_lo + (((kk >> 2) * (128 * 128) + (kk & 3) * 32) >> 4), hi,
                       bk + (((kk >> 2) * (64 * 128) + (kk & 3) * 32) >> 4), hi, idesc_s, kk != 0);
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_f16_ss2(dp_, do_lo + (((kk >> 2) * (128 * 128) + (kk & 3) * 32) >> 4), hi,
                       bv + (((kk >> 2) * (64 * 128) + (kk & 3) * 32) >> 4), hi, idesc_s, kk != 0);
        umma_commit(&sp_full[u]);
      }
      __syncwarp();
    };
    mbar_wait(qdo_full, 0);
    mbar_wait(&k_full[0], 0);
    mbar_wait(&v_full[0], 0);
    tc_fence_after();
    issue_s_dp(0, 0);
    for (int j = 0; j < nkv; ++j) {
      const int u = j & 1, st = j & 1;
      if (j + 1 < nkv) {
        mbar_wait(&k_full[st ^ 1], ((j + 1) >> 1) & 1);
        mbar_wait(&v_full[st ^ 1], ((j + 1) >> 1) & 1);
        // TMEM buffer u^1 was released by ds_full[u^1] of block j-1 (waited below in the previous iteration)
        tc_fence_after();
        issue_s_dp(u ^ 1, st ^ 1);
      }
      mbar_wait(&ds_full[u], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t a = ds_lo + u * (Cfg::kSBytes >> 4), b = kmn_lo + st * (Cfg::kBlkBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16_ss2(tmem_base, a + ((kk * 32) >> 4), hi, b + ((kk * 2048) >> 4), hi, idesc_q, (j > 0) || (kk != 0));
        umma_commit(&kv_empty[st]);
        if (j + 1 == nkv) umma_commit(dq_full);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;  // which 32 of the 64 key columns of a block this thread handles
    const int r = quad * 32 + lane;
    const int row = q0 + r;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const long long bh = (long long)b * p.H + h;
    const float lse2 = row < p.N ? p.lse[bh * p.N + row] * kFaLog2e : INFINITY;
    const float del = row < p.N ? p.delta[bh * p.N + row] : 0.f;
    const float c = p.scale * kFaLog2e;
    const bool drop = p.drop_seed != nullptr;  // dP = mask o (dO V^T) / (1 - p): regenerate the forward mask
    const uint32_t drop_row = drop ? (drop_site_seed(*p.drop_seed, p.drop_site) ^
                                      (((uint32_t)bh * (uint32_t)p.N + (uint32_t)row) * kDropRowMul))
                                   : 0u;
    for (int j = 0; j < nkv; ++j) {
      const int u = j & 1;
      mbar_wait(&sp_full[u], (j >> 1) & 1);
      tc_fence_after();
      uint32_t a0[32], d0[32];
      const uint32_t s_addr = tmem_base + lane_addr + kColS + u * 128 + half * 32;
      tmem_ld_32x32b_x32(s_addr, a0);
      tmem_ld_32x32b_x32(s_addr + 64, d0);
      tc_wait_ld();
      float e[32];
      const int key0 = j * 64 + half * 32;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x0 = fmaf(__uint_as_float(a0[i]), c, -lse2);
        const float p0 = (key0 + i < p.N) ? ((i & 1) ? poly_exp2(x0) : fast_exp2(x0)) : 0.f;  // MUFU / FMA pipes alternate
        float dpv = __uint_as_float(d0[i]);
        if (drop) {
          const uint32_t hsh = drop_mix(drop_row ^ (((uint32_t)(key0 + i) >> 1) * kDropColMul));  // CSE'd per pair
          dpv = (((i & 1) ? (hsh >> 16) : (hsh & 0xffffu)) >= p.drop_thresh16) ? dpv * p.drop_scale : 0.f;
        }
        e[i] = p0 * (dpv - del);
      }
      store_half_row_bf16_sw128(sdS + u * Cfg::kSBytes + r * 128, r, half, e);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full[u]);
    }
    mbar_wait(dq_full, 0);
    tc_fence_after();
    __nv_bfloat16* orow = p.dq + (long long)b * p.qkv_bs + (long long)h * p.qkv_hs + (long long)row * p.qkv_rs;
#pragma unroll 1
    for (int cc = half * (DH / 2); cc < (half + 1) * (DH / 2); cc += 32) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tmem_base + lane_addr + cc, o);
      tc_wait_ld();
      if (row < p.N) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[i]) * p.scale, __uint_as_float(o[i + 1]) * p.scale);
          w.y = pack_bf16x2(__uint_as_float(o[i + 2]) * p.scale, __uint_as_float(o[i + 3]) * p.scale);
          w.z = pack_bf16x2(__uint_as_float(o[i + 4]) * p.scale, __uint_as_float(o[i + 5]) * p.scale);
          w.w = pack_bf16x2(__uint_as_float(o[i + 6]) * p.scale, __uint_as_float(o[i + 7]) * p.scale);
          *reinterpret_cast<uint4*>(orow + cc + i) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ----------------------------------------------- dK, dV -----------------------------------------------
template <int DH>
__global__ void __launch_bounds__(kFaBwdThreads, 1)
fa_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tma_kv128, const __grid_constant__ CUtensorMap tma_q64,
                     const __grid_constant__ CUtensorMap tma_do64, const FaBwdParams p) {
  using Cfg = FaBwdCfg<DH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;                          // [DH/64][128][128B]
  uint8_t* sV = sK + Cfg::kTileBytes;
  uint8_t* sQ = sV + Cfg::kTileBytes;          // [2][DH/64][64][128B]
  uint8_t* sdO = sQ + 2 * Cfg::kBlkBytes;      // [2][DH/64][64][128B]
  uint8_t* sPT = sdO + 2 * Cfg::kBlkBytes;     // [128][128B]
  uint8_t* sdST = sPT + Cfg::kSBytes;          // [128][128B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdST + Cfg::kSBytes);
  uint64_t* kv_full = bars;        // [1]
  uint64_t* q_full = bars + 1;     // [2]
  uint64_t* do_full = bars + 3;    // [2]
  uint64_t* qdo_empty = bars + 5;  // [2]
  uint64_t* sp_full = bars + 7;    // [1] S^T and dP^T complete
  uint64_t* s_free = bars + 8;     // [1] element-wise warps hold S^T / dP^T in registers
  uint64_t* pds_full = bars + 9;   // [1] P^T / dS^T written to smem
  uint64_t* pds_free = bars + 10;  // [1] dV / dK MMAs that read P^T / dS^T retired
  uint64_t* acc_full = bars + 11;  // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* s_lse = reinterpret_cast<float*>(bars + 16);  // [2][64]
  float* s_del = s_lse + 128;                          // [2][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int k0 = blockIdx.x * 128;
  const int nq = (p.N + 63) / 64;
  const int row_base = (int)(b * p.row_bs);
  const int cq = (int)(b * p.col_bs) + p.col_q + h * DH;
  const int ck = (int)(b * p.col_bs) + p.col_k + h * DH;
  const int cv = (int)(b * p.col_bs) + p.col_v + h * DH;
  const int o_row_base = (int)(b * p.o_row_bs);
  const int co = (int)(b * p.o_col_bs) + h * DH;
  constexpr int kColS = 2 * DH;  // TMEM: dV [0,DH) dK [DH,2DH) S^T [2DH,+64) dP^T [2DH+64,+64)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_kv128);
    tma_prefetch_desc(&tma_q64);
    tma_prefetch_desc(&tma_do64);
    mbar_init(kv_full, 1);
    mbar_init(sp_full, 1);
    mbar_init(s_free, 8);  // one arrive per element-wise warp
    mbar_init(pds_full, 8);
    mbar_init(pds_free, 1);
    mbar_init(acc_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&do_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * Cfg::kTileBytes);
#pragma unroll
      for (int c = 0; c < DH / 64; ++c) {
        tma_load_2d(sK + c * (128 * 128), &tma_kv128, kv_full, ck + 64 * c, row_base + k0);
        tma_load_2d(sV + c * (128 * 128), &tma_kv128, kv_full, cv + 64 * c, row_base + k0);
      }
      for (int j = 0; j < nq; ++j) {
        const int st = j & 1;
        mbar_wait(&qdo_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[st], Cfg::kBlkBytes);
#pragma unroll
        for (int c = 0; c < DH / 64; ++c)
          tma_load_2d(sQ + st * Cfg::kBlkBytes + c * (64 * 128), &tma_q64, &q_full[st], cq + 64 * c, row_base + j * 64);
        mbar_expect_tx(&do_full[st], Cfg::kBlkBytes);
#pragma unroll
        for (int c = 0; c < DH / 64; ++c)
          tma_load_2d(sdO + st * Cfg::kBlkBytes + c * (64 * 128), &tma_do64, &do_full[st], co + 64 * c, o_row_base + j * 64);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_a = make_idesc_bf16(128, DH, 0, 1);
    constexpr uint32_t hi = smem_desc_hi_sw128(1024);
    const uint32_t k_lo = smem_desc_lo(smem_u32(sK), 16), v_lo = smem_desc_lo(smem_u32(sV), 16);
    const uint32_t q_lo = smem_desc_lo(smem_u32(sQ), 16), do_lo = smem_desc_lo(smem_u32(sdO), 16);
    const uint32_t qmn_lo = smem_desc_lo(smem_u32(sQ), 64 * 128), domn_lo = smem_desc_lo(smem_u32(sdO), 64 * 128);
    const uint32_t pt_lo = smem_desc_lo(smem_u32(sPT), 16), dst_lo = smem_desc_lo(smem_u32(sdST), 16);
    auto issue_st_dpt = [&](int st) {
      const uint32_t bq = q_lo + st * (Cfg::kBlkBytes >> 4), bo = do_lo + st * (Cfg::kBlkBytes >> 4);
      const uint32_t dst = tmem_base + kColS, ddp = dst + 64;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_f16_ss2(dst, k_lo + (((kk >> 2) * (128 * 128) + (kk & 3) * 32) >> 4), hi,
                       bq + (((kk >> 2) * (64 * 128) + (kk & 3) * 32) >> 4), hi, idesc_s, kk != 0);
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_f16_ss2(ddp, v_lo + (((kk >> 2) * (128 * 128) + (kk & 3) * 32) >> 4), hi,
                       bo + (((kk >> 2) * (64 * 128) + (kk & 3) * 32) >> 4), hi, idesc_s, kk != 0);
        umma_commit(sp_full);
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    mbar_wait(&q_full[0], 0);
    mbar_wait(&do_full[0], 0);
    tc_fence_after();
    issue_st_dpt(0);
    for (int j = 0; j < nq; ++j) {
      const int st = j & 1;
      if (j + 1 < nq) {
        mbar_wait(&q_full[st ^ 1], ((j + 1) >> 1) & 1);
        mbar_wait(&do_full[st ^ 1], ((j + 1) >> 1) & 1);
        mbar_wait(s_free, j & 1);  // S^T / dP^T of block j are in registers: the TMEM columns can be overwritten
        tc_fence_after();
        issue_st_dpt(st ^ 1);
      }
      mbar_wait(pds_full, j & 1);
      tc_fence_after();
      const uint32_t bo = domn_lo + st * (Cfg::kBlkBytes >> 4), bq = qmn_lo + st * (Cfg::kBlkBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // dV += P^T dO
          umma_f16_ss2(tmem_base, pt_lo + ((kk * 32) >> 4), hi, bo + ((kk * 2048) >> 4), hi, idesc_a, (j > 0) || (kk != 0));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // dK += dS^T Q
          umma_f16_ss2(tmem_base + DH, dst_lo + ((kk * 32) >> 4), hi, bq + ((kk * 2048) >> 4), hi, idesc_a, (j > 0) || (kk != 0));
        umma_commit(&qdo_empty[st]);
        umma_commit(pds_free);
        if (j + 1 == nq) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;      // which 32 of the 64 query columns of a block this thread handles
    const int r = quad * 32 + lane;        // key row within the tile
    const int tid = threadIdx.x - 64;      // 0..255 within the element-wise group
    const int krow = k0 + r;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const long long bh = (long long)b * p.H + h;
    const float c = p.scale * kFaLog2e;
    const bool drop = p.drop_seed != nullptr;  // this thread owns key column krow of the mask; queries vary along i
    const uint32_t drop_col = drop ? (drop_site_seed(*p.drop_seed, p.drop_site) ^ (((uint32_t)krow >> 1) * kDropColMul)) : 0u;
    const bool drop_hi = (krow & 1) != 0;
    for (int j = 0; j < nq; ++j) {
      const int u = j & 1;
      if (tid < 128) {  // stage lse / delta of this query block (padded queries: lse = +inf -> P = 0)
        const int qi = j * 64 + (tid & 63);
        if (tid < 64) s_lse[u * 64 + tid] = qi < p.N ? p.lse[bh * p.N + qi] * kFaLog2e : INFINITY;
        else s_del[u * 64 + (tid & 63)] = qi < p.N ? p.delta[bh * p.N + qi] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(sp_full, j & 1);
      tc_fence_after();
      uint32_t a0[32], d0[32];
      const uint32_t s_addr = tmem_base + lane_addr + kColS + half * 32;
      tmem_ld_32x32b_x32(s_addr, a0);
      tmem_ld_32x32b_x32(s_addr + 64, d0);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      float pt[32], ds[32];
      const float* lrow = s_lse + u * 64 + half * 32;
      const float* drow = s_del + u * 64 + half * 32;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x0 = fmaf(__uint_as_float(a0[i]), c, -lrow[i]);
        const float p0 = (i & 1) ? poly_exp2(x0) : fast_exp2(x0);  // MUFU / FMA pipes alternate
        pt[i] = p0;
        ds[i] = __uint_as_float(d0[i]);
      }
      if (drop) {
        // Lanes 2m and 2m+1 own keys of the same column pair, i.e. they need the SAME 32-bit hash for a given query and
        // take different halves of it: each computes the hash for every other query and gets the rest from its partner.
        const uint32_t q_base = (uint32_t)bh * (uint32_t)p.N + (uint32_t)(j * 64 + half * 32);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const uint32_t mine = drop_mix(drop_col ^ ((q_base + (uint32_t)(i + (lane & 1))) * kDropRowMul));
          const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
          const uint32_t h0 = (lane & 1) ? other : mine, h1 = (lane & 1) ? mine : other;  // queries i, i + 1
          const bool k0 = (drop_hi ? (h0 >> 16) : (h0 & 0xffffu)) >= p.drop_thresh16;
          const bool k1 = (drop_hi ? (h1 >> 16) : (h1 & 0xffffu)) >= p.drop_thresh16;
          // dV = (P o mask / (1 - p))^T dO;  dP = mask o (dO V^T) / (1 - p)
          const float pa = pt[i], pb = pt[i + 1];
          pt[i] = k0 ? pa * p.drop_scale : 0.f;
          pt[i + 1] = k1 ? pb * p.drop_scale : 0.f;
          ds[i] = pa * ((k0 ? ds[i] * p.drop_scale : 0.f) - drow[i]);
          ds[i + 1] = pb * ((k1 ? ds[i + 1] * p.drop_scale : 0.f) - drow[i + 1]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) ds[i] = pt[i] * (ds[i] - drow[i]);
      }
      if (j > 0) mbar_wait(pds_free, (j - 1) & 1);  // dV / dK MMAs of block j-1 no longer read the smem tiles
      store_half_row_bf16_sw128(sPT + r * 128, r, half, pt);
      store_half_row_bf16_sw128(sdST + r * 128, r, half, ds);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    __nv_bfloat16* kr = p.dk + (long long)b * p.qkv_bs + (long long)h * p.qkv_hs + (long long)krow * p.qkv_rs;
    __nv_bfloat16* vr = p.dv + (long long)b * p.qkv_bs + (long long)h * p.qkv_hs + (long long)krow * p.qkv_rs;
#pragma unroll 1
    for (int cc = half * (DH / 2); cc < (half + 1) * (DH / 2); cc += 32) {
      uint32_t ov[32], ok[32];
      tmem_ld_32x32b_x32(tmem_base + lane_addr + cc, ov);
      tmem_ld_32x32b_x32(tmem_base + lane_addr + DH + cc, ok);
      tc_wait_ld();
      if (krow < p.N) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(ov[i]), __uint_as_float(ov[i + 1]));
          w.y = pack_bf16x2(__uint_as_float(ov[i + 2]), __uint_as_float(ov[i + 3]));
          w.z = pack_bf16x2(__uint_as_float(ov[i + 4]), __uint_as_float(ov[i + 5]));
          w.w = pack_bf16x2(__uint_as_float(ov[i + 6]), __uint_as_float(ov[i + 7]));
          *reinterpret_cast<uint4*>(vr + cc + i) = w;
          w.x = pack_bf16x2(__uint_as_float(ok[i]) * p.scale, __uint_as_float(ok[i + 1]) * p.scale);
          w.y = pack_bf16x2(__uint_as_float(ok[i + 2]) * p.scale, __uint_as_float(ok[i + 3]) * p.scale);
          w.z = pack_bf16x2(__uint_as_float(ok[i + 4]) * p.scale, __uint_as_float(ok[i + 5]) * p.scale);
          w.w = pack_bf16x2(__uint_as_float(ok[i + 6]) * p.scale, __uint_as_float(ok[i + 7]) * p.scale);
          *reinterpret_cast<uint4*>(kr + cc + i) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

__global__ void __launch_bounds__(256) fa_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                                      float* __restrict__ delta, int B, int H, int N, int DH, long long o_bs,
                                                      long long o_hs, long long o_rs) {
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (long long)B * H * N) return;
  const int i = (int)(warp % N);
  const long long bh = warp / N;
  const long long off = (bh / H) * o_bs + (bh % H) * o_hs + (long long)i * o_rs;
  float s = 0.f;
  for (int d = lane * 2; d < DH; d += 64) {
    const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(o + off + d));
    const float2 g = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dout + off + d));
    s += a.x * g.x + a.y * g.y;
  }
  s = warp_sum(s);
  if (lane == 0) delta[warp] = s;
}

template <int DH>
static int fa_bwd_launch(const AttnParams& a, cudaStream_t stream) {
  using Cfg = FaBwdCfg<DH>;
  const long long E = (long long)a.H * DH;
  if (a.k - a.q != E || a.v - a.q != 2 * E || a.qkv_hs != DH || a.o_hs != DH) return S3D_ERR_UNSUPPORTED;
  if (a.dk - a.dq != E || a.dv - a.dq != 2 * E) return S3D_ERR_UNSUPPORTED;
  FaBwdParams p{};
  long long rows_total, width, o_width;
  if (a.qkv_rs == 3 * E && (a.qkv_bs == (long long)a.N * 3 * E || a.B == 1) && a.o_rs == E &&
      (a.o_bs == (long long)a.N * E || a.B == 1)) {  // timm
    rows_total = (long long)a.B * a.N;
    width = 3 * E;
    o_width = E;
    p.row_bs = a.N;
    p.col_bs = 0;
    p.o_row_bs = a.N;
    p.o_col_bs = 0;
  } else if (a.qkv_bs == 3 * E && a.qkv_rs == (long long)a.B * 3 * E && a.o_bs == E && a.o_rs == (long long)a.B * E) {
    rows_total = a.N;  // sequence-first
    width = (long long)a.B * 3 * E;
    o_width = (long long)a.B * E;
    p.row_bs = 0;
    p.col_bs = 3 * E;
    p.o_row_bs = 0;
    p.o_col_bs = E;
  } else {
    return S3D_ERR_UNSUPPORTED;
  }
  if (a.B > 65535 || a.H > 65535) return S3D_ERR_BAD_SHAPE;
  CUtensorMap t128, t64, d128, d64;
  int rc;
  if ((rc = make_tmap_bf16_2d(&t128, a.q, (uint64_t)width, (uint64_t)rows_total, (uint64_t)a.qkv_rs, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&t64, a.q, (uint64_t)width, (uint64_t)rows_total, (uint64_t)a.qkv_rs, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&d128, a.dout, (uint64_t)o_width, (uint64_t)rows_total, (uint64_t)a.o_rs, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&d64, a.dout, (uint64_t)o_width, (uint64_t)rows_total, (uint64_t)a.o_rs, 64, 64))) return rc;
  p.dq = a.dq;
  p.dk = a.dk;
  p.dv = a.dv;
  p.lse = a.lse;
  p.delta = a.delta;
  p.N = a.N;
  p.H = a.H;
  p.col_q = 0;
  p.col_k = (int)E;
  p.col_v = (int)(2 * E);
  p.qkv_bs = a.qkv_bs;
  p.qkv_hs = a.qkv_hs;
  p.qkv_rs = a.qkv_rs;
  p.scale = a.scale;
  p.drop_seed = a.drop_seed;
  p.drop_site = a.drop_site;
  p.drop_thresh16 = a.drop_thresh16;
  p.drop_scale = a.drop_scale;
  {
    const long long rows = (long long)a.B * a.H * a.N;
    fa_delta_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(a.o, a.dout, a.delta, a.B, a.H, a.N, DH, a.o_bs, a.o_hs, a.o_rs);
    S3D_LAUNCH_OK();
  }
  auto kq = fa_bwd_dq_tc_kernel<DH>;
  auto kkv = fa_bwd_dkv_tc_kernel<DH>;
  static bool attr_set = false;
  if (!attr_set) {
    S3D_CUDA_OK(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    S3D_CUDA_OK(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  dim3 grid((a.N + 127) / 128, a.H, a.B);
  kkv<<<grid, kFaBwdThreads, Cfg::kSmemBytes, stream>>>(t128, t64, d64, p);
  S3D_LAUNCH_OK();
  kq<<<grid, kFaBwdThreads, Cfg::kSmemBytes, stream>>>(t128, t64, d128, p);
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int attn_bwd_tc(const AttnParams& p, int DH, cudaStream_t stream) {
  if (p.B <= 0 || p.H <= 0 || p.N <= 0) return S3D_ERR_BAD_SHAPE;
  switch (DH) {
    case 192: return fa_bwd_launch<192>(p, stream);
    case 64: return fa_bwd_launch<64>(p, stream);
    default: return S3D_ERR_UNSUPPORTED;
  }
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
