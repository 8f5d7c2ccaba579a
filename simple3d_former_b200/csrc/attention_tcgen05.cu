// Flash attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA) for long sequences -- the group_embed layer of the
// reference (nn.TransformerEncoderLayer over S = B*196 = 12544 tokens, dh = E/4, vit_3d_2d_pretrain.py:381,479), which the
// reference evaluates by materialising 15*4 score matrices of S x S fp32 (37.8 GB).
//
// Forward: one CTA owns TWO 128-row query tiles (A, B) of one (batch, head) and streams 64-key K/V tiles through a
// 2-stage TMA ring. Per tile and K/V block:
//     S = Q K^T              tcgen05.mma M128 N64 K16 x DH/16, accumulator in TMEM
//     P = exp2(S*c - m)      softmax warps: tcgen05.ld -> registers -> bf16 -> 128B-swizzled smem
//     O += P V               tcgen05.mma M128 N=DH K16 x 4, V read as an MN-major B operand, O in TMEM
// The two tiles ping-pong: while the softmax warps of tile A work on S_A, the tensor core runs S_B / P_B V. Running
// maxima are updated lazily (O is rescaled only when the max grew by more than 2^8; the row sums absorb the rest exactly).
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = softmax of tile A, warps 6-9 = tile B.
// TMEM (512 columns): O_A [0,DH)  O_B [DH,2DH)  S_A [2DH, 2DH+64)  S_B [2DH+64, 2DH+128).
//
// The element-wise warps are what paces these kernels (ncu, round 1: 16 issued instructions per score element with
// dropout, ALU pipe 120 % and XU pipe 95 % of the MMA time, 20 % of all issued instructions in mbarrier spin loops), so
// their loops are written for instruction count, per pipe:
//   * scale / subtract / row-sum / dS arithmetic on packed fp32 pairs (FFMA2 / FADD2 / FMUL2), 3-input max (FMNMX3);
//   * every exponential on MUFU.EX2 and nothing else on the XU pipe (the former polynomial path used floorf + float->int
//     conversions, which also issue on the XU pipe);
//   * dropout masks from a 5-instruction pair hash whose two 14-bit draws are compared by ONE half2 instruction
//     (common.cuh); the words of the NEXT block are computed after the current block's hand-over, i.e. while the tensor
//     core works, so they are off the S -> P critical path;
//   * mbarrier waits park the warp (try_wait with a suspend-time hint) instead of spinning.
// Head dimensions: any multiple of 16 up to 192 fits the TMEM / shared-memory budget; 48, 64, 96, 192 are instantiated
// (group_embed uses nhead = 4: deit_tiny 48, deit_small 96, deit_base 192). Operand rows are loaded in 64-column TMA
// boxes; for DH % 64 != 0 the tail box carries columns of the neighbouring head that no MMA ever reads.
#include "kernels.h"

#include <stdlib.h>

namespace s3d {

constexpr int kFaBM = 128;   // query rows per tile
constexpr int kFaBN = 64;    // keys per K/V block
constexpr int kFaThreads = 320;
constexpr float kFaLog2e = 1.4426950408889634f;

template <int DH>
struct FaCfg {
  static_assert(DH % 16 == 0 && DH >= 16 && DH <= 192, "head_dim must be a multiple of 16, at most 192");
  static constexpr int kCh = (DH + 63) / 64;              // 64-column (128-byte) chunks per operand row
  static constexpr int kQBytes = kFaBM * kCh * 128;       // per query tile
  static constexpr int kKVBytes = kFaBN * kCh * 128;      // K or V block
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
  uint32_t drop_site, drop_thresh14;
  float drop_scale;           // 1 / (1 - p)
};

__device__ __forceinline__ void tmem_st_32x32b_x4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

// k-step kk (16 bf16 = 32 bytes) inside a [chunks][ROWS][128 B] K-major operand tile, as a descriptor-low-word increment
template <int ROWS>
__device__ __forceinline__ constexpr uint32_t kstep_off(int kk) {
  return (uint32_t)(((kk >> 2) * (ROWS * 128) + (kk & 3) * 32) >> 4);
}

template <int DH, bool DROP>
__global__ void __launch_bounds__(kFaThreads, 1)
fa_fwd_tc_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv, const FaParams p) {
  using Cfg = FaCfg<DH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // [2][kCh][128][128B]
  uint8_t* sK = sQ + 2 * Cfg::kQBytes;                 // [2 stages][kCh][64][128B]
  uint8_t* sV = sK + 2 * Cfg::kKVBytes;                // [2 stages][kCh][64][128B]
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
        for (int c = 0; c < Cfg::kCh; ++c)
          tma_load_2d(sQ + t * Cfg::kQBytes + c * (kFaBM * 128), &tma_q, &q_full[t], cq + 64 * c, row_base + q0 + t * kFaBM);
      }
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], Cfg::kKVBytes);
#pragma unroll
        for (int c = 0; c < Cfg::kCh; ++c)
          tma_load_2d(sK + st * Cfg::kKVBytes + c * (kFaBN * 128), &tma_kv, &k_full[st], ck + 64 * c, row_base + j * kFaBN);
        mbar_expect_tx(&v_full[st], Cfg::kKVBytes);
#pragma unroll
        for (int c = 0; c < Cfg::kCh; ++c)
          tma_load_2d(sV + st * Cfg::kKVBytes + c * (kFaBN * 128), &tma_kv, &v_full[st], cv + 64 * c, row_base + j * kFaBN);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------- MMA issuer ----------------------------------------------
    // The whole warp runs the control flow (waits are warp-uniform); one elected lane issues. Descriptor low words are
    // precomputed so each tcgen05.mma costs a couple of integer adds (at N = 64 an MMA lasts only ~32 cycles, so the
    // issue path must stay far below that).
    constexpr uint32_t idesc_s = make_idesc_bf16(kFaBM, kFaBN, 0, 0);  // S = Q K^T : both K-major
    constexpr uint32_t idesc_o = make_idesc_bf16(kFaBM, DH, 0, 1);     // O = P V   : V is MN-major
    constexpr uint32_t hi = smem_desc_hi_sw128(1024);
    const uint32_t q_lo = smem_desc_lo(smem_u32(sQ), 16), k_lo = smem_desc_lo(smem_u32(sK), 16);
    const uint32_t p_lo = smem_desc_lo(smem_u32(sP), 16), v_lo = smem_desc_lo(smem_u32(sV), kFaBN * 128);
    auto issue_s = [&](int t, int st) {
      const uint32_t a = q_lo + t * (Cfg::kQBytes >> 4), bb = k_lo + st * (Cfg::kKVBytes >> 4);
      const uint32_t d = tmem_base + Cfg::kColS + t * kFaBN;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_f16_ss2(d, a + kstep_off<kFaBM>(kk), hi, bb + kstep_off<kFaBN>(kk), hi, idesc_s, kk != 0);
        umma_commit(&s_full[t]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int t, int st, int j, uint64_t* done0, uint64_t* done1) {
      const uint32_t a = p_lo + t * (Cfg::kPBytes >> 4), bb = v_lo + st * (Cfg::kKVBytes >> 4);
      const uint32_t d = tmem_base + t * DH;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kFaBN / 16; ++kk)
          umma_f16_ss2(d, a + ((kk * 32) >> 4), hi, bb + ((kk * 2048) >> 4), hi, idesc_o, (j > 0) || (kk != 0));
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
    const float2 c2 = make_float2(c, c);
    float m_ref = -INFINITY, l = 0.f;
    // dropout on the attention probabilities (MultiheadAttention(dropout=p) inside nn.TransformerEncoderLayer): P V uses
    // P o mask, the softmax normaliser the full row sum; mask element = (row (b*H + h)*N + query, column key)
    uint32_t mk[DROP ? 32 : 1];  // AND-masks of the 32 packed P pairs of the coming block
    uint32_t ykey = 0, thresh2 = 0;
    if (DROP) {
      ykey = drop_rowkey(drop_site_seed(*p.drop_seed, p.drop_site),
                         (uint32_t)(b * p.H + h) * (uint32_t)p.N + (uint32_t)(q0 + t * kFaBM + r));
      thresh2 = p.drop_thresh14 * 0x00010001u;
#pragma unroll
      for (int i = 0; i < 32; ++i) mk[i] = drop_andmask(drop_word(ykey + (uint32_t)i * kDropColMul), thresh2);
    }
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
      float mx4[4];  // 4 independent chains of 3-input maxima
#pragma unroll
      for (int q = 0; q < 4; ++q) mx4[q] = fmax3(s[q * 16], s[q * 16 + 1], s[q * 16 + 2]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int i = 3; i < 15; i += 2) mx4[q] = fmax3(mx4[q], s[q * 16 + i], s[q * 16 + i + 1]);
        mx4[q] = fmaxf(mx4[q], s[q * 16 + 15]);
      }
      const float mx = fmaxf(fmax3(mx4[0], mx4[1], mx4[2]), mx4[3]);
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
      const float nmc = -m_ref * c;
      const float2 nmc2 = make_float2(nmc, nmc);
      float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 x = ffma2(make_float2(s[ch * 8 + 2 * i], s[ch * 8 + 2 * i + 1]), c2, nmc2);
          const float2 e = make_float2(fast_exp2(x.x), fast_exp2(x.y));
          if (i & 1) sum_b = fadd2(sum_b, e); else sum_a = fadd2(sum_a, e);
          w[i] = pack_bf16x2(e.x, e.y);
          if (DROP) w[i] &= mk[ch * 4 + i];
        }
        *reinterpret_cast<uint4*>(prow + ((ch ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      l += (sum_a.x + sum_a.y) + (sum_b.x + sum_b.y);
      fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      if (DROP) {  // masks of block j + 1, computed while the tensor core runs P V / the next S
        const uint32_t y0 = ykey + (uint32_t)((j + 1) * (kFaBN / 2)) * kDropColMul;
#pragma unroll
        for (int i = 0; i < 32; ++i) mk[i] = drop_andmask(drop_word(y0 + (uint32_t)i * kDropColMul), thresh2);
      }
    }
    // epilogue: O / l -> bf16, lse
    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    const int row = q0 + t * kFaBM + r;
    const float inv = (DROP ? p.drop_scale : 1.0f) / l;
    __nv_bfloat16* orow = p.out + (long long)b * p.o_bs + (long long)h * p.o_hs + (long long)row * p.o_rs;
#pragma unroll 1
    for (int cc = 0; cc < DH; cc += 32) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(o_addr + cc, o);
      tc_wait_ld();
      if (row < p.N) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          if (cc + i < DH) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + cc + i) = u;
          }
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
template <int DH, bool DROP>
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
  p.drop_thresh14 = a.drop_thresh14;
  p.drop_scale = a.drop_scale;
  if ((a.o_rs % 8) || (a.o_hs % 8) || (a.o_bs % 8) || (reinterpret_cast<uintptr_t>(a.out) & 15)) return S3D_ERR_ALIGNMENT;
  auto kern = fa_fwd_tc_kernel<DH, DROP>;
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
// Forward, second generation: ONE 128-row query tile per CTA with every A operand in tensor memory.
//   * Q (bf16 pairs, DH/2 columns) is written to TMEM once by the softmax warps; S = Q K^T runs in TS mode. Shared-memory
//     A operands cost the tensor core ~110-130 cycles per tcgen05.mma whatever N is (measured on the backward kernels:
//     24 SS-mode M128 N64 MMAs per block took ~2700 cycles for 768 cycles of math), so a 64-key S block needs its A
//     operand in tensor memory to run at the 32-cycle rate.
//   * P is written back to TMEM as bf16 pairs (32 columns) by the softmax warps -- no shared-memory tile, no proxy fence --
//     and O += P V runs in TS mode as well (V stays an MN-major B operand in shared memory).
//   * S is single-buffered: the softmax warps pull a block into registers and hand the columns back at once (s_free), so
//     S(j+1) runs on the tensor core while P(j) is computed; the K/V ring has 3-4 stages because shared memory now holds
//     nothing else.
//   * 8 softmax warps: two threads per row (warps w and w+4 share a TMEM lane quadrant), 32 columns each; the row maximum
//     is exchanged through shared memory with a 64-thread named barrier per quadrant, row sums are combined once at the end.
// TMEM: O [0,DH)  S [DH,+64)  P [DH+64,+32)  Q [DH+96,+DH/2)  -- 480 columns at DH = 256, so the timm Block shapes
// (deit_base: 3 heads of 256) run on tcgen05 too. Any N >= 1, head dims 48 / 64 / 96 / 192 / 256.
// ================================================================================================
template <int DH>
struct Fa1Cfg {
  static_assert(DH % 16 == 0 && DH >= 16 && DH <= 256, "head_dim must be a multiple of 16, at most 256");
  static constexpr int kCh = (DH + 63) / 64;
  static constexpr int kBlkBytes = kFaBN * kCh * 128;      // K or V block
  static constexpr int kStages = (2 * 4 * kBlkBytes <= 200 * 1024) ? 4 : 3;
  static constexpr int kSmemBytes = 2 * kStages * kBlkBytes + 1024 + 3072;
  static constexpr int kColS = DH, kColP = DH + 64, kColQ = DH + 96;
  static_assert(DH + 96 + DH / 2 <= 512, "TMEM budget");
};

template <int DH, bool DROP>
__global__ void __launch_bounds__(kFaThreads, 1)
fa_fwd1_tc_kernel(const __grid_constant__ CUtensorMap tma_kv, const __nv_bfloat16* __restrict__ qbase, const FaParams p,
                  long long q_bs, long long q_hs, long long q_rs) {
  using Cfg = Fa1Cfg<DH>;
  constexpr int NST = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;                                // [NST][kCh][64][128B]
  uint8_t* sV = sK + NST * Cfg::kBlkBytes;           // [NST][kCh][64][128B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NST * Cfg::kBlkBytes);
  uint64_t* k_full = bars;                 // [NST]
  uint64_t* v_full = bars + NST;           // [NST]
  uint64_t* kv_empty = bars + 2 * NST;     // [NST]
  uint64_t* s_full = bars + 3 * NST;       // [1]
  uint64_t* s_free = s_full + 1;           // [1] softmax warps hold S in registers
  uint64_t* p_full = s_full + 2;           // [1] P written to tensor memory
  uint64_t* pv_done = s_full + 3;          // [1] P V MMAs of a block retired (P columns and O are idle)
  uint64_t* q_ready = s_full + 4;          // [1]
  uint64_t* o_full = s_full + 5;           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 6);
  float* xch = reinterpret_cast<float*>(s_full + 8);  // [2 parity][2 halves][128 rows] row-maximum exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * kFaBM;
  const int nkv = (p.N + kFaBN - 1) / kFaBN;
  const int row_base = (int)(b * p.row_bs);
  const int ck = (int)(b * p.col_bs) + p.col_k + h * DH;
  const int cv = (int)(b * p.col_bs) + p.col_v + h * DH;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_kv);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 8);
    mbar_init(p_full, 8);
    mbar_init(pv_done, 1);
    mbar_init(q_ready, 8);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // --------------------------------------------- TMA producer ---------------------------------------------
    if (lane == 0) {
      for (int j = 0; j < nkv; ++j) {
        const int st = j % NST;
        mbar_wait(&kv_empty[st], ((j / NST) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], Cfg::kBlkBytes);
#pragma unroll
        for (int c = 0; c < Cfg::kCh; ++c)
          tma_load_2d(sK + st * Cfg::kBlkBytes + c * (kFaBN * 128), &tma_kv, &k_full[st], ck + 64 * c, row_base + j * kFaBN);
        mbar_expect_tx(&v_full[st], Cfg::kBlkBytes);
#pragma unroll
        for (int c = 0; c < Cfg::kCh; ++c)
          tma_load_2d(sV + st * Cfg::kBlkBytes + c * (kFaBN * 128), &tma_kv, &v_full[st], cv + 64 * c, row_base + j * kFaBN);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------- MMA issuer ----------------------------------------------
    constexpr uint32_t idesc_s = make_idesc_bf16(kFaBM, kFaBN, 0, 0);  // S = Q K^T : K block K-major
    constexpr uint32_t idesc_o = make_idesc_bf16(kFaBM, DH, 0, 1);     // O = P V   : V is MN-major
    constexpr uint32_t hi = smem_desc_hi_sw128(1024);
    const uint32_t k_lo = smem_desc_lo(smem_u32(sK), 16), v_lo = smem_desc_lo(smem_u32(sV), kFaBN * 128);
    const uint32_t t_o = tmem_base, t_s = tmem_base + Cfg::kColS, t_p = tmem_base + Cfg::kColP, t_q = tmem_base + Cfg::kColQ;
    auto issue_pv = [&](int j) {
      const int st = j % NST;
      mbar_wait(&v_full[st], (j / NST) & 1);
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      const uint32_t bb = v_lo + st * (Cfg::kBlkBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kFaBN / 16; ++kk)
          umma_f16_ts(t_o, t_p + kk * 8, ((uint64_t)hi << 32) | (bb + ((kk * 2048) >> 4)), idesc_o, (j > 0) || (kk != 0));
        umma_commit(&kv_empty[st]);
        umma_commit(pv_done);
        if (j + 1 == nkv) umma_commit(o_full);
      }
      __syncwarp();
    };
    mbar_wait(q_ready, 0);
    for (int j = 0; j < nkv; ++j) {
      const int st = j % NST;
      mbar_wait(&k_full[st], (j / NST) & 1);
      if (j > 0) mbar_wait(s_free, (j - 1) & 1);
      tc_fence_after();
      const uint32_t bk = k_lo + st * (Cfg::kBlkBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          umma_f16_ts(t_s, t_q + kk * 8, ((uint64_t)hi << 32) | (bk + kstep_off<kFaBN>(kk)), idesc_s, kk != 0);
        umma_commit(s_full);
      }
      __syncwarp();
      if (j > 0) issue_pv(j - 1);
    }
    issue_pv(nkv - 1);
  } else {
    // ----------------------------------------------- softmax -----------------------------------------------
    const int quad = warp & 3;
    const int hf = (warp - 2) >> 2;    // which 32 of the 64 key columns of a block
    const int r = quad * 32 + lane;
    const int row = q0 + r;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + Cfg::kColS + hf * 32;
    const uint32_t p_addr = tmem_base + lane_addr + Cfg::kColP + hf * 16;
    const uint32_t o_addr = tmem_base + lane_addr;
    {  // Q row -> tensor memory (element (row, k) = lane row, column k / 2); rows >= N are zeros
      const __nv_bfloat16* qrow = qbase + (long long)b * q_bs + (long long)h * q_hs + (long long)row * q_rs;
      const uint32_t tq = tmem_base + lane_addr + Cfg::kColQ;
#pragma unroll 4
      for (int g = hf; g < DH / 8; g += 2) {
        uint4 a = make_uint4(0u, 0u, 0u, 0u);
        if (row < p.N) a = *reinterpret_cast<const uint4*>(qrow + g * 8);
        tmem_st_32x32b_x4(tq + g * 4, a.x, a.y, a.z, a.w);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_ready);
    }
    const float c = p.scale * kFaLog2e;
    const float2 c2 = make_float2(c, c);
    float m_ref = -INFINITY, l = 0.f;
    uint32_t mk[DROP ? 16 : 1];  // AND-masks of this thread's 16 packed P pairs of the coming block
    uint32_t ykey = 0, thresh2 = 0;
    if (DROP) {
      ykey = drop_rowkey(drop_site_seed(*p.drop_seed, p.drop_site), (uint32_t)(b * p.H + h) * (uint32_t)p.N + (uint32_t)row) +
             (uint32_t)(hf * 16) * kDropColMul;
      thresh2 = p.drop_thresh14 * 0x00010001u;
#pragma unroll
      for (int i = 0; i < 16; ++i) mk[i] = drop_andmask(drop_word(ykey + (uint32_t)i * kDropColMul), thresh2);
    }
    const int bar_id = 1 + quad;  // named barrier of the two warps that share this lane quadrant
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t v0[32];
      tmem_ld_32x32b_x32(s_addr, v0);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      float s[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(v0[i]);
      const int key0 = j * kFaBN + hf * 32;
      if (j * kFaBN + kFaBN > p.N) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (key0 + i >= p.N) s[i] = -INFINITY;
      }
      float mx4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        mx4[q] = fmax3(s[q * 8], s[q * 8 + 1], s[q * 8 + 2]);
        mx4[q] = fmax3(mx4[q], s[q * 8 + 3], s[q * 8 + 4]);
        mx4[q] = fmax3(mx4[q], s[q * 8 + 5], s[q * 8 + 6]);
        mx4[q] = fmaxf(mx4[q], s[q * 8 + 7]);
      }
      float mx = fmaxf(fmax3(mx4[0], mx4[1], mx4[2]), mx4[3]);
      float* xrow = xch + (j & 1) * 256;
      xrow[hf * 128 + r] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      mx = fmaxf(mx, xrow[(hf ^ 1) * 128 + r]);
      // lazy rescale: keep the reference max unless it grew by more than 2^8 (P stays <= 256, exact in the row sums);
      // both threads of a row see the same mx and take the same decision
      const bool need = (mx - m_ref) * c > 8.0f;
      bool waited = false;
      if (__any_sync(0xffffffffu, need)) {
        const float f = need ? fast_exp2((m_ref - mx) * c) : 1.0f;  // m_ref = -inf -> f = 0 (first block: O is overwritten)
        if (j > 0) {
          mbar_wait(pv_done, (j - 1) & 1);  // O is being accumulated by P V of block j-1 until then
          tc_fence_after();
          waited = true;
#pragma unroll 1
          for (int cc = hf * 16; cc < DH; cc += 32) {  // the two threads of a row take alternate 16-column groups
            uint32_t o[16];
            tmem_ld_32x32b_x16(o_addr + cc, o);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st_32x32b_x16(o_addr + cc, o);
          }
        }
        if (need) { l *= f; m_ref = mx; }
      }
      const float nmc = -m_ref * c;
      const float2 nmc2 = make_float2(nmc, nmc);
      float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 x = ffma2(make_float2(s[2 * i], s[2 * i + 1]), c2, nmc2);
        const float2 e = make_float2(fast_exp2(x.x), fast_exp2(x.y));
        if (i & 1) sum_b = fadd2(sum_b, e); else sum_a = fadd2(sum_a, e);
        w[i] = pack_bf16x2(e.x, e.y);
        if (DROP) w[i] &= mk[i];
      }
      l += (sum_a.x + sum_a.y) + (sum_b.x + sum_b.y);
      if (j > 0 && !waited) {
        mbar_wait(pv_done, (j - 1) & 1);  // P V of block j-1 no longer reads the P columns
        tc_fence_after();
      }
      tmem_st_32x32b_x16(p_addr, w);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (DROP) {  // masks of block j + 1, computed while the tensor core runs
        const uint32_t y0 = ykey + (uint32_t)((j + 1) * (kFaBN / 2)) * kDropColMul;
#pragma unroll
        for (int i = 0; i < 16; ++i) mk[i] = drop_andmask(drop_word(y0 + (uint32_t)i * kDropColMul), thresh2);
      }
    }
    // epilogue: combine the two partial row sums, O / l -> bf16, lse
    float* xrow = xch + (nkv & 1) * 256;
    xrow[hf * 128 + r] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
    l += xrow[(hf ^ 1) * 128 + r];
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = (DROP ? p.drop_scale : 1.0f) / l;
    __nv_bfloat16* orow = p.out + (long long)b * p.o_bs + (long long)h * p.o_hs + (long long)row * p.o_rs;
#pragma unroll 1
    for (int cc = hf * 16; cc < DH; cc += 32) {
      uint32_t o[16];
      tmem_ld_32x32b_x16(o_addr + cc, o);
      tc_wait_ld();
      if (row < p.N) {
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + cc + i) = u;
        }
      }
    }
    if (hf == 0 && p.lse != nullptr && row < p.N) p.lse[((long long)b * p.H + h) * p.N + row] = m_ref * p.scale + logf(l);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int DH, bool DROP>
static int fa_fwd1_launch(const AttnParams& a, cudaStream_t stream) {
  using Cfg = Fa1Cfg<DH>;
  const long long E = (long long)a.H * DH;
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
  CUtensorMap tkv;
  int rc = make_tmap_bf16_2d(&tkv, a.q, (uint64_t)width, (uint64_t)rows_total, (uint64_t)a.qkv_rs, 64, kFaBN);
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
  p.drop_thresh14 = a.drop_thresh14;
  p.drop_scale = a.drop_scale;
  if ((a.o_rs % 8) || (a.o_hs % 8) || (a.o_bs % 8) || (reinterpret_cast<uintptr_t>(a.out) & 15) || (a.qkv_rs % 8) ||
      (a.qkv_bs % 8) || (reinterpret_cast<uintptr_t>(a.q) & 15))
    return S3D_ERR_ALIGNMENT;
  auto kern = fa_fwd1_tc_kernel<DH, DROP>;
  static bool attr_set = false;
  if (!attr_set) {
    S3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  if (a.B > 65535 || a.H > 65535) return S3D_ERR_BAD_SHAPE;
  dim3 grid((a.N + kFaBM - 1) / kFaBM, a.H, a.B);
  kern<<<grid, kFaThreads, Cfg::kSmemBytes, stream>>>(tkv, a.q, p, a.qkv_bs, a.qkv_hs, a.qkv_rs);
  S3D_LAUNCH_OK();
  return S3D_OK;
}

// ================================================================================================
// Backward on tcgen05. Recomputing form (callers without a workspace) = two kernels (no atomics, no round trips through
// HBM); the single-score-pass form built on the dQ kernel is described at FaDqCfg below:
//   dQ   : CTA = 128 query rows; per 64-key block  S = Q K^T, dP = dO V^T (TMEM, double buffered),
//          dS = P o (dP - delta) -> bf16 smem,  dQ += dS K  (K block re-read as an MN-major B operand)
//   dKdV : CTA = 128 key rows; per 64-query block S^T = K Q^T, dP^T = V dO^T,  P^T / dS^T -> bf16 smem,
//          dV += P^T dO,  dK += dS^T Q  (Q / dO blocks re-read as MN-major B operands)
// A single-pass variant (S and dP computed once) would have to hold dK, dV (2 x DH columns), S^T, dP^T and a dQ partial
// in TMEM at once -- 704 columns at dh = 192 against the 512 a CTA owns -- or push 48 KB of fp32 dQ partials per 128 x 64
// tile through L2 reductions (~2.6 TB/s at the target rate); the two-kernel form stays and is made lean instead.
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 element-wise: a TMEM lane (= row) is shared by
// two threads (warps w and w+4 reach the same lane quadrant), each handling 32 of the 64 columns of a block.
// ================================================================================================
struct FaBwdParams {
  __nv_bfloat16 *dq, *dk, *dv;  // outputs, qkv strides
  const float* lse;
  const float* delta;
  int N, H;
  long long row_bs, col_bs;      // qkv 2-D view
  int col_q, col_k, col_v;
  long long o_row_bs, o_col_bs;  // dout 2-D view
  long long o_rs_elems;          // dout row pitch (elements)
  long long qkv_bs, qkv_hs, qkv_rs;
  float scale;
  const uint32_t* drop_seed;  // attention-probability dropout (same mask as the forward kernel)
  uint32_t drop_site, drop_thresh14;
  float drop_scale;
  int dbg;  // S3D_FA_DBG bring-up switches (0 in production)
};

// Element-wise group of the backward kernels: kEwParts threads share one TMEM lane (= tile row), each handling
// 64 / kEwParts columns of a block. Four parts = 16 warps = four per scheduler: the element-wise phase is a chain of
// dependent steps (TMEM load -> exp2 -> mask -> pack -> smem store -> proxy fence) and with two warps per scheduler the
// issue slots sat idle 70 % of the time (ncu: 30 % issue-active, stall_long_sb + stall_wait > 50 %).
constexpr int kEwParts = 4;
constexpr int kEwCols = 64 / kEwParts;            // columns of a 64-wide block per thread
constexpr int kEwWarps = 4 * kEwParts;
constexpr int kEwThreads = 32 * kEwWarps;
constexpr int kFaBwdThreads = 64 + kEwThreads;

template <int DH>
struct FaBwdCfg {
  static constexpr int kCh = (DH + 63) / 64;
  static constexpr int kTileBytes = 128 * kCh * 128;   // 128-row operand tile
  static constexpr int kBlkBytes = 64 * kCh * 128;     // 64-row streamed block
  static constexpr int kSBytes = 128 * 64 * 2;         // bf16 P / dS tile
  static constexpr int kSmemBytes = 2 * kTileBytes + 4 * kBlkBytes + 2 * kSBytes + 1024 + 2048;
};

// kEwCols bf16 (this thread's part of a 128-byte row) into the 128B-swizzled tile: logical 16-byte chunks
// part * kEwCols / 8 ... of row r
__device__ __forceinline__ void store_part_row_sw128(uint8_t* row_base, int r, int part, const uint32_t (&w)[kEwCols / 2]) {
#pragma unroll
  for (int q = 0; q < kEwCols / 8; ++q)
    *reinterpret_cast<uint4*>(row_base + (((part * (kEwCols / 8) + q) ^ (r & 7)) << 4)) =
        make_uint4(w[q * 4], w[q * 4 + 1], w[q * 4 + 2], w[q * 4 + 3]);
}
// streaming store (L2 evict-first: the workspace is read back only after the whole kernel, the K / V operands are what
// should stay L2-resident)
__device__ __forceinline__ void fa_tma_store_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2,
                                                int c3, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;"
               ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[16]) { tmem_ld_32x32b_x16(taddr, v); }
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld_32x32b_x32(taddr, v); }

// ------------------------------------------------ dQ ------------------------------------------------
// The resident operands of this kernel -- the CTA's 128 x DH query tile Q and its dO tile, the A operands of S = Q K^T and
// dP = dO V^T -- live in TENSOR MEMORY (bf16 pairs, DH/2 columns each, written once by the element-wise warps with
// tcgen05.st), not in shared memory:
//   * the tensor core no longer re-reads 2 x 128 x DH bf16 of A operand from shared memory for every 64-key block
//     (ncu, round-2 capture of the smem-A version: tensor-core smem wavefronts 48.6 % of peak at 35 % tensor pipe, i.e. a
//     75 % ceiling; now 88 KB instead of 184 KB per block);
//   * the 96 KB they occupied hold a FOUR-stage K/V ring instead of two stages. With two stages the load of block j+1
//     could only start when dQ(j-1) retired and S(j+1) had to wait for it: load latency sat on the critical path of every
//     block (3037 cycles per block measured against 1152 cycles of MMA work).
// TMEM (2 DH + 128 columns): dQ [0,DH)  S [DH,+64)  dP [DH+64,+64)  Q [DH+128,+DH/2)  dO [DH+128+DH/2,+DH/2).
// S / dP are single-buffered: the element-wise warps pull them into registers and hand the columns back at once
// (sp_free), so S/dP(j+1) runs on the tensor core while dS(j) is computed; dS tiles are double-buffered in smem.
//
// SPILL variant (N >= 128 with a caller workspace, see fa_bwd_spill_launch): this kernel is then the ONLY pass over the score matrix of the
// backward. Besides dQ it writes the two bf16 [query, key] matrices the key-side gradients contract over queries,
//     Pd = P o mask          (dV = Pd^T dO / (1 - p))          dS = P o (mask o dP / (1 - p) - delta)     (dK = dS^T Q scale)
// tile by tile (the dS tile is the A operand of the dQ MMAs and sits in shared memory anyway; the Pd tile gets a buffer
// of its own, paid for with one K/V stage) with one TMA store per tile, and dK / dV become two batched MN-major x MN-major
// GEMMs of gemm_bf16_kernel that read them back once at the HBM rate. 5 GEMM units at (or near) the tensor-memory rate
// instead of 7-8 (dQ 3 + fused dK/dV 4 with shared-memory A operands, or split dK 3 + dV 2), at the price of 8 B of
// workspace traffic per score element (4 B written here at ~3.5 TB/s while the tensor core works, 4 B read by the GEMMs at ~4.9 TB/s).
// MODE 0: recomputing form (dQ only). MODE 1: SPILL (dQ + workspace stores). MODE 2: spill ONLY -- no dQ accumulator, dQ
// is a third GEMM over the spilled dS (its K-major A operand is a panel as it lies): the variant for head_dim 256 (timm
// Blocks of deit_base), whose dQ accumulator + Q + dO + S + dP would need 640 TMEM columns.
template <int DH, int MODE>
struct FaDqCfg {
  static constexpr bool kSpill = MODE > 0, kNoDq = MODE == 2;
  static constexpr int kCh = (DH + 63) / 64;
  static constexpr int kStages = kSpill ? (DH > 192 ? 2 : 3) : 4;
  static constexpr int kBlkBytes = 64 * kCh * 128;     // 64-row streamed block (K or V)
  static constexpr int kSBytes = 128 * 64 * 2;         // bf16 dS (and Pd) tile
  static constexpr int kSmemBytes = 2 * kStages * kBlkBytes + (kSpill ? 4 : 2) * kSBytes + 1024 + 512;
  static constexpr int kColS = kNoDq ? 0 : DH, kColQ = kColS + 128, kColdO = kColQ + DH / 2;
  static_assert(kColdO + DH / 2 <= 512, "TMEM budget");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <int DH, bool DROP, int MODE, int CL>
__global__ void __launch_bounds__(kFaBwdThreads + (MODE > 0 ? 32 : 0), 1)
fa_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tma_kv64, const __grid_constant__ CUtensorMap tma_pd,
                    const __grid_constant__ CUtensorMap tma_ds, const __nv_bfloat16* __restrict__ qbase,
                    const __nv_bfloat16* __restrict__ dobase, const FaBwdParams p) {
  using Cfg = FaDqCfg<DH, MODE>;
  constexpr bool SPILL = Cfg::kSpill, NODQ = Cfg::kNoDq;
  constexpr int NST = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;                                // [NST][kCh][64][128B]
  uint8_t* sV = sK + NST * Cfg::kBlkBytes;           // [NST][kCh][64][128B]
  uint8_t* sdS = sV + NST * Cfg::kBlkBytes;          // [2][128][128B]
  uint8_t* sPd = sdS + 2 * Cfg::kSBytes;             // [2][128][128B] (SPILL only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + (SPILL ? 4 : 2) * Cfg::kSBytes);
  uint64_t* k_full = bars;                 // [NST]
  uint64_t* v_full = bars + NST;           // [NST]
  uint64_t* kv_empty = bars + 2 * NST;     // [NST]
  uint64_t* sp_full = bars + 3 * NST;      // [1] S and dP of a block complete
  uint64_t* sp_free = sp_full + 1;         // [1] element-wise warps hold them in registers
  uint64_t* ds_full = sp_full + 2;         // [2] dS tile written
  uint64_t* ds_free = sp_full + 4;         // [2] dQ MMAs that read the dS tile retired
  uint64_t* a_ready = sp_full + 6;         // [1] Q / dO are in tensor memory
  uint64_t* dq_full = sp_full + 7;         // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sp_full + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * 128;
  const int nkv = (p.N + 63) / 64;
  const int row_base = (int)(b * p.row_bs);
  const int ck = (int)(b * p.col_bs) + p.col_k + h * DH;
  const int cv = (int)(b * p.col_bs) + p.col_v + h * DH;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_kv64);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], CL);  // one (multicast) commit from every CTA of the cluster
    }
    mbar_init(sp_full, 1);
    mbar_init(sp_free, kEwWarps);  // one arrive per element-wise warp
    mbar_init(a_ready, kEwWarps);
    mbar_init(dq_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ds_full[i], kEwWarps);
      mbar_init(&ds_free[i], (SPILL && !NODQ) ? 2 : 1);  // dQ MMAs retired and / or the tile's TMA stores have read it
    }
    fence_barrier_init();
    if (SPILL) {
      tma_prefetch_desc(&tma_pd);
      tma_prefetch_desc(&tma_ds);
    }
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // the peer's barriers are initialised before any multicast can target them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  constexpr uint16_t kMcMask = (uint16_t)((1u << CL) - 1);

  if (warp == 0) {
    if (lane == 0) {
      constexpr int kRows = 64 / CL;  // rows of every 64-row chunk this CTA loads (and multicasts)
      for (int j = 0; j < nkv; ++j) {
        const int st = j % NST;
        mbar_wait(&kv_empty[st], ((j / NST) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], Cfg::kBlkBytes);
#pragma unroll
        for (int c = 0; c < Cfg::kCh; ++c) {
          uint8_t* dst = sK + st * Cfg::kBlkBytes + c * (64 * 128) + crank * kRows * 128;
          if (CL > 1) tma_load_2d_mc(dst, &tma_kv64, &k_full[st], ck + 64 * c, row_base + j * 64 + crank * kRows, kMcMask);
          else tma_load_2d(dst, &tma_kv64, &k_full[st], ck + 64 * c, row_base + j * 64);
        }
        mbar_expect_tx(&v_full[st], Cfg::kBlkBytes);
#pragma unroll
        for (int c = 0; c < Cfg::kCh; ++c) {
          uint8_t* dst = sV + st * Cfg::kBlkBytes + c * (64 * 128) + crank * kRows * 128;
          if (CL > 1) tma_load_2d_mc(dst, &tma_kv64, &v_full[st], cv + 64 * c, row_base + j * 64 + crank * kRows, kMcMask);
          else tma_load_2d(dst, &tma_kv64, &v_full[st], cv + 64 * c, row_base + j * 64);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // warp-converged control flow, one elected lane issues; precomputed descriptor low words (see the forward kernel)
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_q = make_idesc_bf16(128, DH, 0, 1);
    constexpr uint32_t hi = smem_desc_hi_sw128(1024);
    const uint32_t k_lo = smem_desc_lo(smem_u32(sK), 16), v_lo = smem_desc_lo(smem_u32(sV), 16);
    const uint32_t ds_lo = smem_desc_lo(smem_u32(sdS), 16), kmn_lo = smem_desc_lo(smem_u32(sK), 64 * 128);
    const uint32_t t_s = tmem_base + Cfg::kColS, t_dp = t_s + 64;
    const uint32_t t_q = tmem_base + Cfg::kColQ, t_do = tmem_base + Cfg::kColdO;
    auto leader = [&]() { return elect_one(); };
    auto issue_dq = [&](int j) {  // dQ += dS(j) K(j)
      const int u = j & 1, st = j % NST;
      mbar_wait(&ds_full[u], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t a = ds_lo + u * (Cfg::kSBytes >> 4), bb = kmn_lo + st * (Cfg::kBlkBytes >> 4);
      if (leader()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16_ss2(tmem_base, a + ((kk * 32) >> 4), hi, bb + ((kk * 2048) >> 4), hi, idesc_q, (j > 0) || (kk != 0));
        if (CL > 1) umma_commit_mc(&kv_empty[st], kMcMask);
        else umma_commit(&kv_empty[st]);
        umma_commit(&ds_free[u]);
        if (j + 1 == nkv) umma_commit(dq_full);
      }
      __syncwarp();
    };
    mbar_wait(a_ready, 0);
    for (int j = 0; j < nkv; ++j) {
      const int st = j % NST;
      mbar_wait(&k_full[st], (j / NST) & 1);
      mbar_wait(&v_full[st], (j / NST) & 1);
      if (j > 0) mbar_wait(sp_free, (j - 1) & 1);  // S / dP of block j-1 are in registers: the columns can be overwritten
      tc_fence_after();
      const uint32_t bk = k_lo + st * (Cfg::kBlkBytes >> 4), bv = v_lo + st * (Cfg::kBlkBytes >> 4);
      if (leader()) {
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)  // S = Q K^T, A from tensor memory (8 columns per 16-element k-step)
          umma_f16_ts(t_s, t_q + kk * 8, ((uint64_t)hi << 32) | (bk + kstep_off<64>(kk)), idesc_s, kk != 0);
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)  // dP = dO V^T
          umma_f16_ts(t_dp, t_do + kk * 8, ((uint64_t)hi << 32) | (bv + kstep_off<64>(kk)), idesc_s, kk != 0);
        umma_commit(sp_full);
        if (NODQ) {  // no dQ MMAs will read this K / V block: the slot is free once S / dP retire
          if (CL > 1) umma_commit_mc(&kv_empty[st], kMcMask);
          else umma_commit(&kv_empty[st]);
        }
      }
      __syncwarp();
      if (!NODQ && j > 0) issue_dq(j - 1);
    }
    if (!NODQ) issue_dq(nkv - 1);
  } else if (SPILL && warp == 2 + kEwWarps) {
    // ---- store warp: one TMA store per finished Pd / dS tile (the element-wise warps fenced their writes for the async
    // proxy before arriving on ds_full); the tile is handed back once the store has read it AND the dQ MMAs retired
    if (lane == 0) {
      const int bh_i = b * p.H + h;
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      for (int j = 0; j < nkv; ++j) {
        const int u = j & 1;
        mbar_wait(&ds_full[u], (j >> 1) & 1);
        if (!(p.dbg & 1)) fa_tma_store_4d(&tma_pd, sPd + u * Cfg::kSBytes, 0, q0, j, bh_i, policy);
        if (!(p.dbg & 2)) fa_tma_store_4d(&tma_ds, sdS + u * Cfg::kSBytes, 0, q0, j, bh_i, policy);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(&ds_free[u]);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;  // which kEwCols of the 64 key columns of a block this thread handles
    const int r = quad * 32 + lane;
    const int row = q0 + r;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const long long bh = (long long)b * p.H + h;
    // ---- Q / dO rows -> tensor memory (bf16 pairs; element (row, k) = lane row, column k / 2). Rows >= N are zeros.
    {
      const __nv_bfloat16* qrow = qbase + (long long)b * p.qkv_bs + (long long)h * p.qkv_hs + (long long)row * p.qkv_rs;
      const __nv_bfloat16* orow = dobase + (long long)(b * p.o_row_bs + row) * p.o_rs_elems + b * p.o_col_bs + h * DH;
      const uint32_t tq = tmem_base + lane_addr + Cfg::kColQ, tdo = tmem_base + lane_addr + Cfg::kColdO;
#pragma unroll 4
      for (int g = part; g < DH / 8; g += kEwParts) {  // 8 bf16 = 16 bytes = 4 columns per step
        uint4 a = make_uint4(0u, 0u, 0u, 0u), d = make_uint4(0u, 0u, 0u, 0u);
        if (row < p.N) {
          a = *reinterpret_cast<const uint4*>(qrow + g * 8);
          d = *reinterpret_cast<const uint4*>(orow + g * 8);
        }
        tmem_st_32x32b_x4(tq + g * 4, a.x, a.y, a.z, a.w);
        tmem_st_32x32b_x4(tdo + g * 4, d.x, d.y, d.z, d.w);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
    }
    // rows >= N (last query tile): lse = +inf -> P = 0 -> dS = 0; they are never written back.
    // keys >= N (last block): dS is zeroed explicitly below -- in the timm layout the K / V rows behind a batch's last key
    // are the next batch's keys, not TMA zero fill.
    const float nlse = row < p.N ? -p.lse[bh * p.N + row] * kFaLog2e : -INFINITY;
    const float ndel = row < p.N ? -p.delta[bh * p.N + row] : 0.f;
    const float c = p.scale * kFaLog2e;
    const float2 c2 = make_float2(c, c), nlse2 = make_float2(nlse, nlse), ndel2 = make_float2(ndel, ndel);
    const float ks = DROP ? p.drop_scale : 1.0f;  // dP = mask o (dO V^T) / (1 - p): regenerate the forward mask
    const float2 ks2 = make_float2(ks, ks);
    uint32_t z[DROP ? kEwCols / 2 : 1];
    uint32_t ykey = 0, thresh2 = 0;
    if (DROP) {
      ykey = drop_rowkey(drop_site_seed(*p.drop_seed, p.drop_site), (uint32_t)bh * (uint32_t)p.N + (uint32_t)row) +
             (uint32_t)(part * (kEwCols / 2)) * kDropColMul;
      thresh2 = p.drop_thresh14 * 0x00010001u;
#pragma unroll
      for (int i = 0; i < kEwCols / 2; ++i) z[i] = drop_word(ykey + (uint32_t)i * kDropColMul);
    }
    for (int j = 0; j < nkv; ++j) {
      const int u = j & 1;
      mbar_wait(sp_full, j & 1);
      tc_fence_after();
      uint32_t a0[kEwCols], d0[kEwCols];
      const uint32_t s_addr = tmem_base + lane_addr + Cfg::kColS + part * kEwCols;
      tmem_ld_cols(s_addr, a0);
      tmem_ld_cols(s_addr + 64, d0);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sp_free);
      uint32_t w[kEwCols / 2];
      uint32_t wp[kEwCols / 2];  // (dead code when !SPILL)
#pragma unroll
      for (int i = 0; i < kEwCols / 2; ++i) {
        const float2 x = ffma2(make_float2(__uint_as_float(a0[2 * i]), __uint_as_float(a0[2 * i + 1])), c2, nlse2);
        const float2 pr = make_float2(fast_exp2(x.x), fast_exp2(x.y));
        float2 dp = make_float2(__uint_as_float(d0[2 * i]), __uint_as_float(d0[2 * i + 1]));
        if (DROP) drop_zero2(dp.x, dp.y, z[i], thresh2);
        const float2 e = fmul2(pr, ffma2(dp, ks2, ndel2));
        w[i] = pack_bf16x2(e.x, e.y);
        if (SPILL) {
          wp[i] = pack_bf16x2(pr.x, pr.y);
          if (DROP) wp[i] &= drop_andmask(z[i], thresh2);  // Pd = P o mask (the 1 / (1 - p) is the dV GEMM's alpha)
        }
      }
      if (j * 64 + 64 > p.N) {  // last, partial key block
        const int key0 = j * 64 + part * kEwCols;
#pragma unroll
        for (int i = 0; i < kEwCols / 2; ++i) {
          if (key0 + 2 * i >= p.N) w[i] = 0u;
          else if (key0 + 2 * i + 1 >= p.N) w[i] &= 0xffffu;
          if (SPILL) {
            if (key0 + 2 * i >= p.N) wp[i] = 0u;
            else if (key0 + 2 * i + 1 >= p.N) wp[i] &= 0xffffu;
          }
        }
      }
      if (j >= 2) mbar_wait(&ds_free[u], ((j >> 1) - 1) & 1);  // dQ MMAs of block j-2 no longer read this dS tile
      store_part_row_sw128(sdS + u * Cfg::kSBytes + r * 128, r, part, w);
      if constexpr (SPILL) { if (!(p.dbg & 4)) store_part_row_sw128(sPd + u * Cfg::kSBytes + r * 128, r, part, wp); }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full[u]);
      if (DROP) {  // mask words of block j + 1, computed while the tensor core runs
        const uint32_t y0 = ykey + (uint32_t)((j + 1) * 32) * kDropColMul;
#pragma unroll
        for (int i = 0; i < kEwCols / 2; ++i) z[i] = drop_word(y0 + (uint32_t)i * kDropColMul);
      }
    }
    if constexpr (!NODQ) {  // (MODE 2: dQ comes from the third GEMM)
    mbar_wait(dq_full, 0);
    tc_fence_after();
    __nv_bfloat16* orow = p.dq + (long long)b * p.qkv_bs + (long long)h * p.qkv_hs + (long long)row * p.qkv_rs;
    // the threads of a row split the DH columns in 16-column groups: group g -> part g % kEwParts
#pragma unroll 1
    for (int cc = part * 16; cc < DH; cc += 16 * kEwParts) {
      uint32_t o[16];
      tmem_ld_32x32b_x16(tmem_base + lane_addr + cc, o);
      tc_wait_ld();
      if (row < p.N) {
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          uint4 wv;
          wv.x = pack_bf16x2(__uint_as_float(o[i]) * p.scale, __uint_as_float(o[i + 1]) * p.scale);
          wv.y = pack_bf16x2(__uint_as_float(o[i + 2]) * p.scale, __uint_as_float(o[i + 3]) * p.scale);
          wv.z = pack_bf16x2(__uint_as_float(o[i + 4]) * p.scale, __uint_as_float(o[i + 5]) * p.scale);
          wv.w = pack_bf16x2(__uint_as_float(o[i + 6]) * p.scale, __uint_as_float(o[i + 7]) * p.scale);
          *reinterpret_cast<uint4*>(orow + cc + i) = wv;
        }
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no CTA exits while its peer may still multicast into it / signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ----------------------------------------------- dK, dV -----------------------------------------------
template <int DH, bool DROP>
__global__ void __launch_bounds__(kFaBwdThreads, 1)
fa_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tma_kv128, const __grid_constant__ CUtensorMap tma_q64,
                     const __grid_constant__ CUtensorMap tma_do64, const FaBwdParams p) {
  using Cfg = FaBwdCfg<DH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;                          // [kCh][128][128B]
  uint8_t* sV = sK + Cfg::kTileBytes;
  uint8_t* sQ = sV + Cfg::kTileBytes;          // [2][kCh][64][128B]
  uint8_t* sdO = sQ + 2 * Cfg::kBlkBytes;      // [2][kCh][64][128B]
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
  float* s_nlse = reinterpret_cast<float*>(bars + 16);          // [2][64]  -lse * log2(e) of the block's queries
  float* s_ndel = s_nlse + 128;                                 // [2][64]  -delta
  uint32_t* s_rk = reinterpret_cast<uint32_t*>(s_ndel + 128);   // [2][64]  dropout row keys

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
    mbar_init(s_free, kEwWarps);  // one arrive per element-wise warp
    mbar_init(pds_full, kEwWarps);
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
      for (int c = 0; c < Cfg::kCh; ++c) {
        tma_load_2d(sK + c * (128 * 128), &tma_kv128, kv_full, ck + 64 * c, row_base + k0);
        tma_load_2d(sV + c * (128 * 128), &tma_kv128, kv_full, cv + 64 * c, row_base + k0);
      }
      for (int j = 0; j < nq; ++j) {
        const int st = j & 1;
        mbar_wait(&qdo_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[st], Cfg::kBlkBytes);
#pragma unroll
        for (int c = 0; c < Cfg::kCh; ++c)
          tma_load_2d(sQ + st * Cfg::kBlkBytes + c * (64 * 128), &tma_q64, &q_full[st], cq + 64 * c, row_base + j * 64);
        mbar_expect_tx(&do_full[st], Cfg::kBlkBytes);
#pragma unroll
        for (int c = 0; c < Cfg::kCh; ++c)
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
        // the two accumulation chains are interleaved: consecutive tcgen05.mma into the SAME accumulator serialise on the
        // shared-memory operand fetch, MMAs into different accumulators overlap it
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) {
          umma_f16_ss2(dst, k_lo + kstep_off<128>(kk), hi, bq + kstep_off<64>(kk), hi, idesc_s, kk != 0);
          umma_f16_ss2(ddp, v_lo + kstep_off<128>(kk), hi, bo + kstep_off<64>(kk), hi, idesc_s, kk != 0);
        }
        umma_commit(sp_full);
      }
      __syncwarp();
    };
    auto issue_dv_dk = [&](int j) {
      const int st = j & 1;
      const uint32_t bo = domn_lo + st * (Cfg::kBlkBytes >> 4), bq = qmn_lo + st * (Cfg::kBlkBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {  // dV += P^T dO and dK += dS^T Q, interleaved (independent accumulators)
          umma_f16_ss2(tmem_base, pt_lo + ((kk * 32) >> 4), hi, bo + ((kk * 2048) >> 4), hi, idesc_a, (j > 0) || (kk != 0));
          umma_f16_ss2(tmem_base + DH, dst_lo + ((kk * 32) >> 4), hi, bq + ((kk * 2048) >> 4), hi, idesc_a, (j > 0) || (kk != 0));
        }
        umma_commit(&qdo_empty[st]);
        umma_commit(pds_free);
        if (j + 1 == nq) umma_commit(acc_full);
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    mbar_wait(&q_full[0], 0);
    mbar_wait(&do_full[0], 0);
    tc_fence_after();
    issue_st_dpt(0);
    // Dynamic issue order. Two kinds of work are pending: S^T/dP^T of block `ns` (needs its Q/dO stage loaded and the
    // element-wise warps to have pulled block ns-1 out of the single S^T/dP^T TMEM buffer) and dV/dK of block `nd` (needs
    // P^T/dS^T of block nd in smem). With a fixed order the issuer blocks on whichever comes first in program order -- the
    // round-2 capture showed it waiting for the load of block j+1 (which can only start when dV/dK(j-1) retire: two stages)
    // while dV/dK(j) was ready, 3506 cycles per block for 1536 cycles of MMA work. Whatever is ready is issued first.
    int ns = 1, nd = 0;
    while (nd < nq) {
      bool progressed = false;
      if (ns < nq && ns <= nd + 1) {
        const int st = ns & 1;
        const uint32_t ph = (ns >> 1) & 1;
        if (mbar_try_wait(&q_full[st], ph) && mbar_try_wait(&do_full[st], ph) && mbar_try_wait(s_free, (ns - 1) & 1)) {
          tc_fence_after();
          issue_st_dpt(st);
          ++ns;
          progressed = true;
        }
      }
      if (!progressed && mbar_try_wait(pds_full, nd & 1)) {
        tc_fence_after();
        issue_dv_dk(nd);
        ++nd;
        progressed = true;
      }
      if (!progressed) __nanosleep(32);
    }
  } else {
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;      // which kEwCols of the 64 query columns of a block this thread handles
    const int r = quad * 32 + lane;        // key row within the tile
    const int tid = threadIdx.x - 64;      // 0..255 within the element-wise group
    const int krow = k0 + r;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const long long bh = (long long)b * p.H + h;
    const float c = p.scale * kFaLog2e;
    const float2 c2 = make_float2(c, c);
    const float ks = DROP ? p.drop_scale : 1.0f;
    const float2 ks2 = make_float2(ks, ks);
    // dropout: this thread owns key column krow of the mask, the block's queries are the mask rows. Word of (query, krow >> 1)
    // = drop_word(rowkey(query) + colterm); the draw is its low / high half for an even / odd key: one byte permute picks
    // the halves of two consecutive queries into one word for the half2 comparison.
    const uint32_t site_seed = DROP ? drop_site_seed(*p.drop_seed, p.drop_site) : 0u;
    const uint32_t colterm = ((uint32_t)krow >> 1) * kDropColMul;
    const uint32_t sel = (krow & 1) ? 0x7632u : 0x5410u;
    const uint32_t thresh2 = p.drop_thresh14 * 0x00010001u;
    uint32_t z[DROP ? kEwCols / 2 : 1];  // per query pair (2i, 2i+1) of this thread's queries
    // lse / delta / row keys of a query block (padded queries: lse = +inf -> P = 0): global loads two blocks ahead into a
    // register, shared-memory store one block ahead (see the dK kernel below)
    auto stage_load = [&](int j) -> float {
      const int qi = j * 64 + (tid & 63);
      if (tid < 64) return (j < nq && qi < p.N) ? -p.lse[bh * p.N + qi] * kFaLog2e : -INFINITY;
      if (tid < 128) return (j < nq && qi < p.N) ? -p.delta[bh * p.N + qi] : 0.f;
      return 0.f;
    };
    auto stage_store = [&](int j, float val) {
      const int u = j & 1;
      if (tid < 64) s_nlse[u * 64 + tid] = val;
      else if (tid < 128) s_ndel[u * 64 + (tid & 63)] = val;
      else if (DROP && tid < 192)
        s_rk[u * 64 + (tid & 63)] = drop_rowkey(site_seed, (uint32_t)bh * (uint32_t)p.N + (uint32_t)(j * 64 + (tid & 63)));
    };
    auto make_words = [&](int j) {
      const uint4* rk4 = reinterpret_cast<const uint4*>(s_rk + (j & 1) * 64 + part * kEwCols);
#pragma unroll
      for (int g = 0; g < kEwCols / 4; ++g) {
        const uint4 rk = rk4[g];
        const uint32_t z0 = drop_word(rk.x + colterm), z1 = drop_word(rk.y + colterm);
        const uint32_t z2 = drop_word(rk.z + colterm), z3 = drop_word(rk.w + colterm);
        z[2 * g] = __byte_perm(z0, z1, sel);
        z[2 * g + 1] = __byte_perm(z2, z3, sel);
      }
    };
    stage_store(0, stage_load(0));
    float pend = stage_load(1);
    asm volatile("bar.sync 1, %0;" ::"n"(kEwThreads) : "memory");
    if (DROP) make_words(0);
    for (int j = 0; j < nq; ++j) {
      const int u = j & 1;
      if (j + 1 < nq) stage_store(j + 1, pend);  // buffer u^1: its last readers finished block j-1 before the barrier of j-1
      pend = stage_load(j + 2);
      mbar_wait(sp_full, j & 1);
      tc_fence_after();
      uint32_t a0[kEwCols], d0[kEwCols];
      const uint32_t s_addr = tmem_base + lane_addr + kColS + part * kEwCols;
      tmem_ld_cols(s_addr, a0);
      tmem_ld_cols(s_addr + 64, d0);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      uint32_t wp[kEwCols / 2], wd[kEwCols / 2];
      const float2* lrow = reinterpret_cast<const float2*>(s_nlse + u * 64 + part * kEwCols);
      const float2* drow = reinterpret_cast<const float2*>(s_ndel + u * 64 + part * kEwCols);
#pragma unroll
      for (int i = 0; i < kEwCols / 2; ++i) {
        const float2 x = ffma2(make_float2(__uint_as_float(a0[2 * i]), __uint_as_float(a0[2 * i + 1])), c2, lrow[i]);
        const float2 pr = make_float2(fast_exp2(x.x), fast_exp2(x.y));
        float2 dp = make_float2(__uint_as_float(d0[2 * i]), __uint_as_float(d0[2 * i + 1]));
        float2 pk = pr;
        if (DROP) {
          // dV = (P o mask / (1 - p))^T dO  (the 1 / (1 - p) is applied to the dV accumulator at the end);
          // dP = mask o (dO V^T) / (1 - p)
          drop_zero2(dp.x, dp.y, z[i], thresh2);
          drop_zero2(pk.x, pk.y, z[i], thresh2);
        }
        const float2 e = fmul2(pr, ffma2(dp, ks2, drow[i]));
        wp[i] = pack_bf16x2(pk.x, pk.y);
        wd[i] = pack_bf16x2(e.x, e.y);
      }
      if (j > 0) mbar_wait(pds_free, (j - 1) & 1);  // dV / dK MMAs of block j-1 no longer read the smem tiles
      store_part_row_sw128(sPT + r * 128, r, part, wp);
      store_part_row_sw128(sdST + r * 128, r, part, wd);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      asm volatile("bar.sync 1, %0;" ::"n"(kEwThreads) : "memory");  // staging of block j + 1 is complete / buffer u may be refilled
      if (DROP && j + 1 < nq) make_words(j + 1);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    __nv_bfloat16* kr = p.dk + (long long)b * p.qkv_bs + (long long)h * p.qkv_hs + (long long)krow * p.qkv_rs;
    __nv_bfloat16* vr = p.dv + (long long)b * p.qkv_bs + (long long)h * p.qkv_hs + (long long)krow * p.qkv_rs;
#pragma unroll 1
    for (int cc = part * 16; cc < DH; cc += 16 * kEwParts) {
      uint32_t ov[16], ok[16];
      tmem_ld_32x32b_x16(tmem_base + lane_addr + cc, ov);
      tmem_ld_32x32b_x16(tmem_base + lane_addr + DH + cc, ok);
      tc_wait_ld();
      if (krow < p.N) {
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          uint4 wv;
          wv.x = pack_bf16x2(__uint_as_float(ov[i]) * ks, __uint_as_float(ov[i + 1]) * ks);
          wv.y = pack_bf16x2(__uint_as_float(ov[i + 2]) * ks, __uint_as_float(ov[i + 3]) * ks);
          wv.z = pack_bf16x2(__uint_as_float(ov[i + 4]) * ks, __uint_as_float(ov[i + 5]) * ks);
          wv.w = pack_bf16x2(__uint_as_float(ov[i + 6]) * ks, __uint_as_float(ov[i + 7]) * ks);
          *reinterpret_cast<uint4*>(vr + cc + i) = wv;
          wv.x = pack_bf16x2(__uint_as_float(ok[i]) * p.scale, __uint_as_float(ok[i + 1]) * p.scale);
          wv.y = pack_bf16x2(__uint_as_float(ok[i + 2]) * p.scale, __uint_as_float(ok[i + 3]) * p.scale);
          wv.z = pack_bf16x2(__uint_as_float(ok[i + 4]) * p.scale, __uint_as_float(ok[i + 5]) * p.scale);
          wv.w = pack_bf16x2(__uint_as_float(ok[i + 6]) * p.scale, __uint_as_float(ok[i + 7]) * p.scale);
          *reinterpret_cast<uint4*>(kr + cc + i) = wv;
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

// ------------------------------------------- dK and dV, split -------------------------------------------
// The fused dK/dV kernel above needs 2 DH accumulator columns + S^T + dP^T = 512 TMEM columns at DH = 192, which leaves
// no room for its resident A operands (the K and V tiles): they stay in shared memory and every one of its 24 S^T / dP^T
// MMAs per block pays the ~110-cycle shared-memory A fetch (3800 cycles per block measured for 1536 cycles of math).
// Split in two, every A operand fits tensor memory:
//   dK kernel: K, V tiles in TMEM;  S^T = K Q^T, dP^T = V dO^T (TS mode), dS^T -> smem, dK += dS^T Q   (3 GEMM units)
//   dV kernel: K tile in TMEM;      S^T = K Q^T (TS), P^T -> TMEM as bf16 pairs, dV += P^T dO (TS)     (2 GEMM units)
// One more GEMM unit than the fused form (S^T twice), but all of it at the tensor-memory rate; both kernels stream the
// Q / dO blocks through a 4-stage ring. The dK kernel is the dQ kernel with rows and columns exchanged (softmax statistics
// and dropout rows belong to the streamed queries, i.e. to the COLUMNS of a block, and are staged in shared memory).
template <int DH>
struct FaDkCfg {
  static constexpr int kCh = (DH + 63) / 64;
  static constexpr int kStages = 4;
  static constexpr int kBlkBytes = 64 * kCh * 128;     // 64-row streamed block (Q or dO)
  static constexpr int kSBytes = 128 * 64 * 2;         // bf16 dS^T tile
  static constexpr int kSmemBytes = 2 * kStages * kBlkBytes + 2 * kSBytes + 1024 + 2048;  // 232448 at DH = 192: the opt-in limit
  static constexpr int kColS = DH, kColK = DH + 128, kColV = DH + 128 + DH / 2;
  static_assert(2 * DH + 128 <= 512, "TMEM budget");
};

__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// rows of a resident 128-row operand tile -> tensor memory (bf16 pairs; element (row, k) = lane row, column k / 2)
template <int DH>
__device__ __forceinline__ void fill_tmem_rows(uint32_t taddr, const __nv_bfloat16* grow, bool valid, int part) {
#pragma unroll 4
  for (int g = part; g < DH / 8; g += kEwParts) {  // 8 bf16 = 16 bytes = 4 columns per step
    uint4 a = make_uint4(0u, 0u, 0u, 0u);
    if (valid) a = *reinterpret_cast<const uint4*>(grow + g * 8);
    tmem_st_32x32b_x4(taddr + g * 4, a.x, a.y, a.z, a.w);
  }
}

template <int DH, bool DROP>
__global__ void __launch_bounds__(kFaBwdThreads, 1)
fa_bwd_dk_tc_kernel(const __grid_constant__ CUtensorMap tma_q64, const __grid_constant__ CUtensorMap tma_do64,
                    const __nv_bfloat16* __restrict__ kbase, const __nv_bfloat16* __restrict__ vbase, const FaBwdParams p) {
  using Cfg = FaDkCfg<DH>;
  constexpr int NST = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                // [NST][kCh][64][128B]
  uint8_t* sdO = sQ + NST * Cfg::kBlkBytes;          // [NST][kCh][64][128B]
  uint8_t* sdST = sdO + NST * Cfg::kBlkBytes;        // [2][128][128B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdST + 2 * Cfg::kSBytes);
  uint64_t* qdo_full = bars;               // [NST]
  uint64_t* qdo_empty = bars + NST;        // [NST]
  uint64_t* sp_full = bars + 2 * NST;      // [1]
  uint64_t* s_free = sp_full + 1;          // [1]
  uint64_t* ds_full = sp_full + 2;         // [2]
  uint64_t* ds_free = sp_full + 4;         // [2]
  uint64_t* a_ready = sp_full + 6;         // [1]
  uint64_t* acc_full = sp_full + 7;        // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sp_full + 8);
  float* s_nlse = reinterpret_cast<float*>(sp_full + 10);       // [2][64]  -lse * log2(e) of the block's queries
  float* s_ndel = s_nlse + 128;                                 // [2][64]  -delta
  uint32_t* s_rk = reinterpret_cast<uint32_t*>(s_ndel + 128);   // [2][64]  dropout row keys

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int k0 = blockIdx.x * 128;
  const int nq = (p.N + 63) / 64;
  const int row_base = (int)(b * p.row_bs);
  const int cq = (int)(b * p.col_bs) + p.col_q + h * DH;
  const int o_row_base = (int)(b * p.o_row_bs);
  const int co = (int)(b * p.o_col_bs) + h * DH;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q64);
    tma_prefetch_desc(&tma_do64);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    mbar_init(sp_full, 1);
    mbar_init(s_free, kEwWarps);
    mbar_init(a_ready, kEwWarps);
    mbar_init(acc_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ds_full[i], kEwWarps);
      mbar_init(&ds_free[i], 1);
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
      for (int j = 0; j < nq; ++j) {
        const int st = j % NST;
        mbar_wait(&qdo_empty[st], ((j / NST) & 1) ^ 1);
        mbar_expect_tx(&qdo_full[st], 2 * Cfg::kBlkBytes);
#pragma unroll
        for (int c = 0; c < Cfg::kCh; ++c) {
          tma_load_2d(sQ + st * Cfg::kBlkBytes + c * (64 * 128), &tma_q64, &qdo_full[st], cq + 64 * c, row_base + j * 64);
          tma_load_2d(sdO + st * Cfg::kBlkBytes + c * (64 * 128), &tma_do64, &qdo_full[st], co + 64 * c, o_row_base + j * 64);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_a = make_idesc_bf16(128, DH, 0, 1);
    constexpr uint32_t hi = smem_desc_hi_sw128(1024);
    const uint32_t q_lo = smem_desc_lo(smem_u32(sQ), 16), do_lo = smem_desc_lo(smem_u32(sdO), 16);
    const uint32_t qmn_lo = smem_desc_lo(smem_u32(sQ), 64 * 128), dst_lo = smem_desc_lo(smem_u32(sdST), 16);
    const uint32_t t_s = tmem_base + Cfg::kColS, t_dp = t_s + 64;
    const uint32_t t_k = tmem_base + Cfg::kColK, t_v = tmem_base + Cfg::kColV;
    auto issue_dk = [&](int j) {  // dK += dS^T(j) Q(j)
      const int u = j & 1, st = j % NST;
      mbar_wait(&ds_full[u], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t a = dst_lo + u * (Cfg::kSBytes >> 4), bb = qmn_lo + st * (Cfg::kBlkBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16_ss2(tmem_base, a + ((kk * 32) >> 4), hi, bb + ((kk * 2048) >> 4), hi, idesc_a, (j > 0) || (kk != 0));
        umma_commit(&qdo_empty[st]);
        umma_commit(&ds_free[u]);
        if (j + 1 == nq) umma_commit(acc_full);
      }
      __syncwarp();
    };
    mbar_wait(a_ready, 0);
    for (int j = 0; j < nq; ++j) {
      const int st = j % NST;
      mbar_wait(&qdo_full[st], (j / NST) & 1);
      if (j > 0) mbar_wait(s_free, (j - 1) & 1);
      tc_fence_after();
      const uint32_t bq = q_lo + st * (Cfg::kBlkBytes >> 4), bo = do_lo + st * (Cfg::kBlkBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)  // S^T = K Q^T
          umma_f16_ts(t_s, t_k + kk * 8, ((uint64_t)hi << 32) | (bq + kstep_off<64>(kk)), idesc_s, kk != 0);
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)  // dP^T = V dO^T
          umma_f16_ts(t_dp, t_v + kk * 8, ((uint64_t)hi << 32) | (bo + kstep_off<64>(kk)), idesc_s, kk != 0);
        umma_commit(sp_full);
      }
      __syncwarp();
      if (j > 0) issue_dk(j - 1);
    }
    issue_dk(nq - 1);
  } else {
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;      // which kEwCols of the 64 query columns of a block this thread handles
    const int r = quad * 32 + lane;        // key row within the tile
    const int tid = threadIdx.x - 64;
    const int krow = k0 + r;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const long long bh = (long long)b * p.H + h;
    {
      const long long off = (long long)b * p.qkv_bs + (long long)h * p.qkv_hs + (long long)krow * p.qkv_rs;
      fill_tmem_rows<DH>(tmem_base + lane_addr + Cfg::kColK, kbase + off, krow < p.N, part);
      fill_tmem_rows<DH>(tmem_base + lane_addr + Cfg::kColV, vbase + off, krow < p.N, part);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
    }
    const float c = p.scale * kFaLog2e;
    const float2 c2 = make_float2(c, c);
    const float ks = DROP ? p.drop_scale : 1.0f;
    const float2 ks2 = make_float2(ks, ks);
    const uint32_t site_seed = DROP ? drop_site_seed(*p.drop_seed, p.drop_site) : 0u;
    const uint32_t colterm = ((uint32_t)krow >> 1) * kDropColMul;
    const uint32_t sel = (krow & 1) ? 0x7632u : 0x5410u;
    const uint32_t thresh2 = p.drop_thresh14 * 0x00010001u;
    uint32_t z[DROP ? kEwCols / 2 : 1];
    // Per-query statistics of a block (-lse, -delta, dropout row key) are staged in shared memory by the first 192
    // threads. The global loads are issued TWO blocks ahead into a register and stored one block ahead, so their latency
    // never sits in front of the block barrier (with a plain load-then-store the staging warps stalled ~700 cycles at the
    // top of every block and everybody waited for them).
    auto stage_load = [&](int j) -> float {
      const int qi = j * 64 + (tid & 63);
      if (tid < 64) return (j < nq && qi < p.N) ? -p.lse[bh * p.N + qi] * kFaLog2e : -INFINITY;
      if (tid < 128) return (j < nq && qi < p.N) ? -p.delta[bh * p.N + qi] : 0.f;
      return 0.f;
    };
    auto stage_store = [&](int j, float val) {
      const int u = j & 1;
      if (tid < 64) s_nlse[u * 64 + tid] = val;
      else if (tid < 128) s_ndel[u * 64 + (tid & 63)] = val;
      else if (DROP && tid < 192)
        s_rk[u * 64 + (tid & 63)] = drop_rowkey(site_seed, (uint32_t)bh * (uint32_t)p.N + (uint32_t)(j * 64 + (tid & 63)));
    };
    auto make_words = [&](int j) {
      const uint4* rk4 = reinterpret_cast<const uint4*>(s_rk + (j & 1) * 64 + part * kEwCols);
#pragma unroll
      for (int g = 0; g < kEwCols / 4; ++g) {
        const uint4 rk = rk4[g];
        const uint32_t z0 = drop_word(rk.x + colterm), z1 = drop_word(rk.y + colterm);
        const uint32_t z2 = drop_word(rk.z + colterm), z3 = drop_word(rk.w + colterm);
        z[2 * g] = __byte_perm(z0, z1, sel);
        z[2 * g + 1] = __byte_perm(z2, z3, sel);
      }
    };
    stage_store(0, stage_load(0));
    float pend = stage_load(1);
    asm volatile("bar.sync 1, %0;" ::"n"(kEwThreads) : "memory");
    if (DROP) make_words(0);
    for (int j = 0; j < nq; ++j) {
      const int u = j & 1;
      if (j + 1 < nq) stage_store(j + 1, pend);
      pend = stage_load(j + 2);
      mbar_wait(sp_full, j & 1);
      tc_fence_after();
      uint32_t a0[kEwCols], d0[kEwCols];
      const uint32_t s_addr = tmem_base + lane_addr + Cfg::kColS + part * kEwCols;
      tmem_ld_cols(s_addr, a0);
      tmem_ld_cols(s_addr + 64, d0);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      uint32_t wd[kEwCols / 2];
      const float2* lrow = reinterpret_cast<const float2*>(s_nlse + u * 64 + part * kEwCols);
      const float2* drow = reinterpret_cast<const float2*>(s_ndel + u * 64 + part * kEwCols);
#pragma unroll
      for (int i = 0; i < kEwCols / 2; ++i) {
        const float2 x = ffma2(make_float2(__uint_as_float(a0[2 * i]), __uint_as_float(a0[2 * i + 1])), c2, lrow[i]);
        const float2 pr = make_float2(fast_exp2(x.x), fast_exp2(x.y));
        float2 dp = make_float2(__uint_as_float(d0[2 * i]), __uint_as_float(d0[2 * i + 1]));
        if (DROP) drop_zero2(dp.x, dp.y, z[i], thresh2);
        const float2 e = fmul2(pr, ffma2(dp, ks2, drow[i]));
        wd[i] = pack_bf16x2(e.x, e.y);
      }
      if (j >= 2) mbar_wait(&ds_free[u], ((j >> 1) - 1) & 1);
      store_part_row_sw128(sdST + u * Cfg::kSBytes + r * 128, r, part, wd);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full[u]);
      asm volatile("bar.sync 1, %0;" ::"n"(kEwThreads) : "memory");  // staging of block j + 1 complete / buffer u reusable
      if (DROP && j + 1 < nq) make_words(j + 1);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    __nv_bfloat16* kr = p.dk + (long long)b * p.qkv_bs + (long long)h * p.qkv_hs + (long long)krow * p.qkv_rs;
#pragma unroll 1
    for (int cc = part * 16; cc < DH; cc += 16 * kEwParts) {
      uint32_t ok[16];
      tmem_ld_32x32b_x16(tmem_base + lane_addr + cc, ok);
      tc_wait_ld();
      if (krow < p.N) {
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          uint4 wv;
          wv.x = pack_bf16x2(__uint_as_float(ok[i]) * p.scale, __uint_as_float(ok[i + 1]) * p.scale);
          wv.y = pack_bf16x2(__uint_as_float(ok[i + 2]) * p.scale, __uint_as_float(ok[i + 3]) * p.scale);
          wv.z = pack_bf16x2(__uint_as_float(ok[i + 4]) * p.scale, __uint_as_float(ok[i + 5]) * p.scale);
          wv.w = pack_bf16x2(__uint_as_float(ok[i + 6]) * p.scale, __uint_as_float(ok[i + 7]) * p.scale);
          *reinterpret_cast<uint4*>(kr + cc + i) = wv;
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

template <int DH>
struct FaDvCfg {
  static constexpr int kCh = (DH + 63) / 64;
  static constexpr int kStages = 4;
  static constexpr int kBlkBytes = 64 * kCh * 128;
  static constexpr int kSmemBytes = 2 * kStages * kBlkBytes + 1024 + 2048;
  static constexpr int kColS = DH, kColP = DH + 64, kColK = DH + 96;
};

template <int DH, bool DROP>
__global__ void __launch_bounds__(kFaBwdThreads, 1)
fa_bwd_dv_tc_kernel(const __grid_constant__ CUtensorMap tma_q64, const __grid_constant__ CUtensorMap tma_do64,
                    const __nv_bfloat16* __restrict__ kbase, const FaBwdParams p) {
  using Cfg = FaDvCfg<DH>;
  constexpr int NST = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                // [NST][kCh][64][128B]
  uint8_t* sdO = sQ + NST * Cfg::kBlkBytes;          // [NST][kCh][64][128B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdO + NST * Cfg::kBlkBytes);
  uint64_t* q_full = bars;                 // [NST]
  uint64_t* do_full = bars + NST;          // [NST]
  uint64_t* qdo_empty = bars + 2 * NST;    // [NST]
  uint64_t* s_full = bars + 3 * NST;       // [1]
  uint64_t* s_free = s_full + 1;           // [1]
  uint64_t* p_full = s_full + 2;           // [1] P^T written to tensor memory
  uint64_t* pv_done = s_full + 3;          // [1] dV MMAs of a block retired
  uint64_t* a_ready = s_full + 4;          // [1]
  uint64_t* acc_full = s_full + 5;         // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 6);
  float* s_nlse = reinterpret_cast<float*>(s_full + 8);         // [2][64]
  uint32_t* s_rk = reinterpret_cast<uint32_t*>(s_nlse + 128);   // [2][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int k0 = blockIdx.x * 128;
  const int nq = (p.N + 63) / 64;
  const int row_base = (int)(b * p.row_bs);
  const int cq = (int)(b * p.col_bs) + p.col_q + h * DH;
  const int o_row_base = (int)(b * p.o_row_bs);
  const int co = (int)(b * p.o_col_bs) + h * DH;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q64);
    tma_prefetch_desc(&tma_do64);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&do_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, kEwWarps);
    mbar_init(p_full, kEwWarps);
    mbar_init(pv_done, 1);
    mbar_init(a_ready, kEwWarps);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int j = 0; j < nq; ++j) {
        const int st = j % NST;
        mbar_wait(&qdo_empty[st], ((j / NST) & 1) ^ 1);
        mbar_expect_tx(&q_full[st], Cfg::kBlkBytes);
#pragma unroll
        for (int c = 0; c < Cfg::kCh; ++c)
          tma_load_2d(sQ + st * Cfg::kBlkBytes + c * (64 * 128), &tma_q64, &q_full[st], cq + 64 * c, row_base + j * 64);
        mbar_expect_tx(&do_full[st], Cfg::kBlkBytes);
#pragma unroll
        for (int c = 0; c < Cfg::kCh; ++c)
          tma_load_2d(sdO + st * Cfg::kBlkBytes + c * (64 * 128), &tma_do64, &do_full[st], co + 64 * c, o_row_base + j * 64);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_a = make_idesc_bf16(128, DH, 0, 1);
    constexpr uint32_t hi = smem_desc_hi_sw128(1024);
    const uint32_t q_lo = smem_desc_lo(smem_u32(sQ), 16), domn_lo = smem_desc_lo(smem_u32(sdO), 64 * 128);
    const uint32_t t_s = tmem_base + Cfg::kColS, t_p = tmem_base + Cfg::kColP, t_k = tmem_base + Cfg::kColK;
    auto issue_dv = [&](int j) {  // dV += P^T(j) dO(j), A from tensor memory
      const int st = j % NST;
      mbar_wait(&do_full[st], (j / NST) & 1);
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      const uint32_t bb = domn_lo + st * (Cfg::kBlkBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16_ts(tmem_base, t_p + kk * 8, ((uint64_t)hi << 32) | (bb + ((kk * 2048) >> 4)), idesc_a, (j > 0) || (kk != 0));
        umma_commit(&qdo_empty[st]);
        umma_commit(pv_done);
        if (j + 1 == nq) umma_commit(acc_full);
      }
      __syncwarp();
    };
    mbar_wait(a_ready, 0);
    for (int j = 0; j < nq; ++j) {
      const int st = j % NST;
      mbar_wait(&q_full[st], (j / NST) & 1);
      if (j > 0) mbar_wait(s_free, (j - 1) & 1);
      tc_fence_after();
      const uint32_t bq = q_lo + st * (Cfg::kBlkBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)  // S^T = K Q^T
          umma_f16_ts(t_s, t_k + kk * 8, ((uint64_t)hi << 32) | (bq + kstep_off<64>(kk)), idesc_s, kk != 0);
        umma_commit(s_full);
      }
      __syncwarp();
      if (j > 0) issue_dv(j - 1);
    }
    issue_dv(nq - 1);
  } else {
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const int tid = threadIdx.x - 64;
    const int krow = k0 + r;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const long long bh = (long long)b * p.H + h;
    {
      const long long off = (long long)b * p.qkv_bs + (long long)h * p.qkv_hs + (long long)krow * p.qkv_rs;
      fill_tmem_rows<DH>(tmem_base + lane_addr + Cfg::kColK, kbase + off, krow < p.N, part);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
    }
    const float c = p.scale * kFaLog2e;
    const float2 c2 = make_float2(c, c);
    const uint32_t site_seed = DROP ? drop_site_seed(*p.drop_seed, p.drop_site) : 0u;
    const uint32_t colterm = ((uint32_t)krow >> 1) * kDropColMul;
    const uint32_t sel = (krow & 1) ? 0x7632u : 0x5410u;
    const uint32_t thresh2 = p.drop_thresh14 * 0x00010001u;
    uint32_t mk[DROP ? kEwCols / 2 : 1];  // AND-masks of the packed P^T pairs (two consecutive queries of this key)
    auto stage_load = [&](int j) -> float {  // issued two blocks ahead (see the dK kernel)
      const int qi = j * 64 + (tid & 63);
      return (tid < 64 && j < nq && qi < p.N) ? -p.lse[bh * p.N + qi] * kFaLog2e : -INFINITY;
    };
    auto stage_store = [&](int j, float val) {
      const int u = j & 1;
      if (tid < 64) s_nlse[u * 64 + tid] = val;
      else if (DROP && tid < 128)
        s_rk[u * 64 + (tid & 63)] = drop_rowkey(site_seed, (uint32_t)bh * (uint32_t)p.N + (uint32_t)(j * 64 + (tid & 63)));
    };
    auto make_masks = [&](int j) {
      const uint4* rk4 = reinterpret_cast<const uint4*>(s_rk + (j & 1) * 64 + part * kEwCols);
#pragma unroll
      for (int g = 0; g < kEwCols / 4; ++g) {
        const uint4 rk = rk4[g];
        const uint32_t z0 = drop_word(rk.x + colterm), z1 = drop_word(rk.y + colterm);
        const uint32_t z2 = drop_word(rk.z + colterm), z3 = drop_word(rk.w + colterm);
        mk[2 * g] = drop_andmask(__byte_perm(z0, z1, sel), thresh2);
        mk[2 * g + 1] = drop_andmask(__byte_perm(z2, z3, sel), thresh2);
      }
    };
    stage_store(0, stage_load(0));
    float pend = stage_load(1);
    asm volatile("bar.sync 1, %0;" ::"n"(kEwThreads) : "memory");
    if (DROP) make_masks(0);
    const uint32_t p_addr = tmem_base + lane_addr + Cfg::kColP + part * (kEwCols / 2);
    for (int j = 0; j < nq; ++j) {
      const int u = j & 1;
      if (j + 1 < nq) stage_store(j + 1, pend);
      pend = stage_load(j + 2);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t a0[kEwCols];
      tmem_ld_cols(tmem_base + lane_addr + Cfg::kColS + part * kEwCols, a0);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      uint32_t wp[kEwCols / 2];
      const float2* lrow = reinterpret_cast<const float2*>(s_nlse + u * 64 + part * kEwCols);
#pragma unroll
      for (int i = 0; i < kEwCols / 2; ++i) {
        const float2 x = ffma2(make_float2(__uint_as_float(a0[2 * i]), __uint_as_float(a0[2 * i + 1])), c2, lrow[i]);
        wp[i] = pack_bf16x2(fast_exp2(x.x), fast_exp2(x.y));
        if (DROP) wp[i] &= mk[i];  // dV = (P o mask / (1 - p))^T dO; the 1 / (1 - p) is applied to the accumulator at the end
      }
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);  // dV MMAs of block j-1 no longer read the P^T columns
        tc_fence_after();
      }
      static_assert(kEwCols / 2 == 8, "P^T store shape");
      tmem_st_32x32b_x8(p_addr, wp);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      asm volatile("bar.sync 1, %0;" ::"n"(kEwThreads) : "memory");
      if (DROP && j + 1 < nq) make_masks(j + 1);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const float ks = DROP ? p.drop_scale : 1.0f;
    __nv_bfloat16* vr = p.dv + (long long)b * p.qkv_bs + (long long)h * p.qkv_hs + (long long)krow * p.qkv_rs;
#pragma unroll 1
    for (int cc = part * 16; cc < DH; cc += 16 * kEwParts) {
      uint32_t ov[16];
      tmem_ld_32x32b_x16(tmem_base + lane_addr + cc, ov);
      tc_wait_ld();
      if (krow < p.N) {
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          uint4 wv;
          wv.x = pack_bf16x2(__uint_as_float(ov[i]) * ks, __uint_as_float(ov[i + 1]) * ks);
          wv.y = pack_bf16x2(__uint_as_float(ov[i + 2]) * ks, __uint_as_float(ov[i + 3]) * ks);
          wv.z = pack_bf16x2(__uint_as_float(ov[i + 4]) * ks, __uint_as_float(ov[i + 5]) * ks);
          wv.w = pack_bf16x2(__uint_as_float(ov[i + 6]) * ks, __uint_as_float(ov[i + 7]) * ks);
          *reinterpret_cast<uint4*>(vr + cc + i) = wv;
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

// Layout parse (timm [B, N, 3, H, dh] or sequence-first [S, Nb, 3E]), parameter block, alignment checks and the delta
// kernel: shared by every form of the backward.
struct FaBwdSetup {
  FaBwdParams p;
  long long rows_total, width, o_width;
};
static int fa_bwd_setup(const AttnParams& a, int DH, FaBwdSetup& st, cudaStream_t stream) {
  const long long E = (long long)a.H * DH;
  if (a.k - a.q != E || a.v - a.q != 2 * E || a.qkv_hs != DH || a.o_hs != DH) return S3D_ERR_UNSUPPORTED;
  if (a.dk - a.dq != E || a.dv - a.dq != 2 * E) return S3D_ERR_UNSUPPORTED;
  FaBwdParams& p = st.p;
  p = FaBwdParams{};
  if (a.qkv_rs == 3 * E && (a.qkv_bs == (long long)a.N * 3 * E || a.B == 1) && a.o_rs == E &&
      (a.o_bs == (long long)a.N * E || a.B == 1)) {  // timm
    st.rows_total = (long long)a.B * a.N;
    st.width = 3 * E;
    st.o_width = E;
    p.row_bs = a.N;
    p.col_bs = 0;
    p.o_row_bs = a.N;
    p.o_col_bs = 0;
  } else if (a.qkv_bs == 3 * E && a.qkv_rs == (long long)a.B * 3 * E && a.o_bs == E && a.o_rs == (long long)a.B * E) {
    st.rows_total = a.N;  // sequence-first
    st.width = (long long)a.B * 3 * E;
    st.o_width = (long long)a.B * E;
    p.row_bs = 0;
    p.col_bs = 3 * E;
    p.o_row_bs = 0;
    p.o_col_bs = E;
  } else {
    return S3D_ERR_UNSUPPORTED;
  }
  if (a.B > 65535 || a.H > 65535) return S3D_ERR_BAD_SHAPE;
  if ((a.qkv_rs % 8) || (a.qkv_hs % 8) || (a.qkv_bs % 8) || (a.o_rs % 8) || (reinterpret_cast<uintptr_t>(a.q) & 15) ||
      (reinterpret_cast<uintptr_t>(a.dout) & 15))
    return S3D_ERR_ALIGNMENT;  // 16-byte row loads of Q / dO
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
  p.drop_thresh14 = a.drop_thresh14;
  p.drop_scale = a.drop_scale;
  p.o_rs_elems = a.o_rs;
  { const char* v = getenv("S3D_FA_DBG"); p.dbg = v == nullptr ? 0 : atoi(v); }
  const long long rows = (long long)a.B * a.H * a.N;
  fa_delta_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(a.o, a.dout, a.delta, a.B, a.H, a.N, DH, a.o_bs, a.o_hs, a.o_rs);
  S3D_LAUNCH_OK();
  return S3D_OK;
}

static bool fa_spill_applies(const AttnParams& a, int DH) {
  const long long need = attn_bwd_workspace_bytes(a.B, a.H, a.N, DH);
  return a.workspace != nullptr && need > 0 && a.workspace_bytes >= need && (a.qkv_bs % DH) == 0 && (a.o_bs % DH) == 0;
}

// ---- single score pass: the dQ kernel spills Pd / dS (bf16 64-key panels) into the caller's workspace, dV / dK (and, in
// MODE 2, dQ) are batched GEMMs over them. Measured on the group_embed shape: DESIGN.md 4.4, profiles/ncu_table.json.
template <int DH, bool DROP, int MODE>
static int fa_bwd_spill_launch(const AttnParams& a, const FaBwdSetup& st, cudaStream_t stream) {
  using QCfg = FaDqCfg<DH, MODE>;
  const FaBwdParams& p = st.p;
  int rc;
  CUtensorMap t64;
  if ((rc = make_tmap_bf16_2d(&t64, a.q, (uint64_t)st.width, (uint64_t)st.rows_total, (uint64_t)a.qkv_rs, 64, 64))) return rc;
  dim3 grid((a.N + 127) / 128, a.H, a.B);
  const long long npad = (a.N + 63) / 64 * 64;
  const long long BH = (long long)a.B * a.H;
  // workspace layout: 64-key panels [B*H][Npad / 64][N queries][64 keys] -- a 128 x 64 tile store and a GEMM operand box
  // are contiguous runs in HBM (row-major [N, Npad] matrices wrote 128-byte pieces 25 KB apart: 3 TB/s)
  __nv_bfloat16* pd = reinterpret_cast<__nv_bfloat16*>(a.workspace);
  __nv_bfloat16* ds = pd + ((2 * BH * a.N * npad + 1023) / 1024 * 1024) / 2;
  CUtensorMap tpd, tds;
  if ((rc = make_tmap_bf16_panel(&tpd, pd, (uint64_t)a.N, (uint64_t)(npad / 64), (uint64_t)BH, 128, 1))) return rc;
  if ((rc = make_tmap_bf16_panel(&tds, ds, (uint64_t)a.N, (uint64_t)(npad / 64), (uint64_t)BH, 128, 1))) return rc;
  // S3D_FA_CL=2: CTA pairs sharing the K / V blocks by multicast. Measured equal (10.9 ms both ways on the group_embed
  // shape): this kernel is bound by its element-wise warps and, with the stores on, by ~3.5 TB/s of HBM writes -- not by
  // the L2 -> SM operand traffic the pairs halve. Off by default.
  static const bool no_cluster = []() { const char* v = getenv("S3D_FA_CL"); return v == nullptr || v[0] != '2'; }();
  if (no_cluster) {
    auto kqs = fa_bwd_dq_tc_kernel<DH, DROP, MODE, 1>;
    static bool attr3_set = false;
    if (!attr3_set) {
      S3D_CUDA_OK(cudaFuncSetAttribute(kqs, cudaFuncAttributeMaxDynamicSharedMemorySize, QCfg::kSmemBytes));
      attr3_set = true;
    }
    kqs<<<grid, kFaBwdThreads + 32, QCfg::kSmemBytes, stream>>>(t64, tpd, tds, a.q, a.dout, p);
    S3D_LAUNCH_OK();
  } else {
    auto kqs = fa_bwd_dq_tc_kernel<DH, DROP, MODE, 2>;
    static bool attr4_set = false;
    if (!attr4_set) {
      S3D_CUDA_OK(cudaFuncSetAttribute(kqs, cudaFuncAttributeMaxDynamicSharedMemorySize, QCfg::kSmemBytes));
      attr4_set = true;
    }
    CUtensorMap t32;  // half-chunk boxes: every CTA of a pair loads 32 of the 64 rows and multicasts them
    if ((rc = make_tmap_bf16_2d(&t32, a.q, (uint64_t)st.width, (uint64_t)st.rows_total, (uint64_t)a.qkv_rs, 64, 32))) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((grid.x + 1) / 2 * 2, grid.y, grid.z);  // an odd last tile gets an idle partner (all rows >= N)
    cfg.blockDim = dim3(kFaBwdThreads + 32, 1, 1);
    cfg.dynamicSmemBytes = QCfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    S3D_CUDA_OK(cudaLaunchKernelEx(&cfg, kqs, t32, tpd, tds, a.q, a.dout, p));
  }
  // dV[b,h] = Pd[b,h]^T dO[b,h] / (1 - p);  dK[b,h] = dS[b,h]^T Q[b,h] * scale: A = workspace matrix read MN-major
  // (m = key contiguous), B = dO / Q read MN-major (n = head channel contiguous), contraction over the queries.
  // A (batch, head) slice of qkv / dout sits at element offset b * batch_stride + h * DH = (b * batch_stride / DH + h) * DH.
  GemmArgs g{};
  g.a_mn = 1;
  g.b_mn = 1;
  g.batch = (int)BH;
  g.batch_stride_b = DH;
  g.batch_inner = a.H;
  g.bmul_a = a.H;
  g.p.M = a.N;
  g.p.N = DH;
  g.p.K = a.N;
  g.p.out_fp32 = 0;
  g.p.epilogue = EPI_NONE;
  g.p.batched = 1;
  g.p.a_panel = 1;
  g.p.batch_stride_d = DH;
  g.p.ldd = a.qkv_rs;
  g.bmul_d = (int)(a.qkv_bs / DH);
  // dV
  g.A = pd;
  g.B = a.dout;
  g.ldb = a.o_rs;
  g.bmul_b = (int)(a.o_bs / DH);
  g.p.D = a.dv;
  g.p.alpha = DROP ? a.drop_scale : 1.0f;
  if ((rc = gemm_bf16(g, stream))) return rc;
  // dK
  g.A = ds;
  g.B = a.q;
  g.ldb = a.qkv_rs;
  g.bmul_b = (int)(a.qkv_bs / DH);
  g.p.D = a.dk;
  g.p.alpha = a.scale;
  if ((rc = gemm_bf16(g, stream))) return rc;
  if (MODE == 2) {
    // dQ[b,h] = dS[b,h] K[b,h] * scale: A = dS read K-major straight from its panels (panel = 64-key k-block), B = K read
    // MN-major, contraction over the keys (keys >= N hold zeros in dS and are zero-filled in K by the tensor map)
    g.a_mn = 0;
    g.p.a_panel = 2;
    g.B = a.k;
    g.p.D = a.dq;
    return gemm_bf16(g, stream);
  }
  return S3D_OK;
}

template <int DH, bool DROP>
static int fa_bwd_launch(const AttnParams& a, cudaStream_t stream) {
  using Cfg = FaBwdCfg<DH>;
  FaBwdSetup st;
  int rc;
  if ((rc = fa_bwd_setup(a, DH, st, stream))) return rc;
  if (fa_spill_applies(a, DH)) return fa_bwd_spill_launch<DH, DROP, 1>(a, st, stream);
  const FaBwdParams& p = st.p;
  const long long rows_total = st.rows_total, width = st.width, o_width = st.o_width;
  CUtensorMap t128, t64, d64;
  if ((rc = make_tmap_bf16_2d(&t128, a.q, (uint64_t)width, (uint64_t)rows_total, (uint64_t)a.qkv_rs, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&t64, a.q, (uint64_t)width, (uint64_t)rows_total, (uint64_t)a.qkv_rs, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&d64, a.dout, (uint64_t)o_width, (uint64_t)rows_total, (uint64_t)a.o_rs, 64, 64))) return rc;
  auto kq = fa_bwd_dq_tc_kernel<DH, DROP, 0, 1>;
  auto kkv = fa_bwd_dkv_tc_kernel<DH, DROP>;
  static bool attr_set = false;
  if (!attr_set) {
    S3D_CUDA_OK(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, FaDqCfg<DH, 0>::kSmemBytes));
    S3D_CUDA_OK(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  dim3 grid((a.N + 127) / 128, a.H, a.B);
  // Split dK / dV kernels (all A operands in tensor memory) or the fused dK/dV kernel (K, V tiles in shared memory)?
  // Measured on the group_embed shape (B = 15, H = 4, S = 12544, dh = 192), dK + dV:
  //     without dropout   split 7.8 + 5.8 = 13.6 ms (tensor pipe 71 % / 62 %)    fused 15.4 ms (41 %)
  //     dropout p = 0.1   split 11.4 + 9.3 = 20.7 ms                             fused 16.2 ms
  // The split form runs the mask hash twice (once per kernel, one hash word per ELEMENT in the key-row orientation) and
  // is then bound by its element-wise warps; the fused form hides the hash behind its shared-memory operand fetches.
  // S3D_FA_DKV=fused / split forces one of them.
  static const int forced = []() { const char* v = getenv("S3D_FA_DKV"); return v == nullptr ? 0 : (v[0] == 'f' ? 1 : (v[0] == 's' ? 2 : 0)); }();
  const bool fused = forced == 1 || (forced == 0 && DROP);
  if (fused) {
    kkv<<<grid, kFaBwdThreads, Cfg::kSmemBytes, stream>>>(t128, t64, d64, p);
    S3D_LAUNCH_OK();
  } else {
    auto kdk = fa_bwd_dk_tc_kernel<DH, DROP>;
    auto kdv = fa_bwd_dv_tc_kernel<DH, DROP>;
    static bool attr2_set = false;
    if (!attr2_set) {
      S3D_CUDA_OK(cudaFuncSetAttribute(kdk, cudaFuncAttributeMaxDynamicSharedMemorySize, FaDkCfg<DH>::kSmemBytes));
      S3D_CUDA_OK(cudaFuncSetAttribute(kdv, cudaFuncAttributeMaxDynamicSharedMemorySize, FaDvCfg<DH>::kSmemBytes));
      attr2_set = true;
    }
    kdk<<<grid, kFaBwdThreads, FaDkCfg<DH>::kSmemBytes, stream>>>(t64, d64, a.k, a.v, p);
    S3D_LAUNCH_OK();
    kdv<<<grid, kFaBwdThreads, FaDvCfg<DH>::kSmemBytes, stream>>>(t64, d64, a.k, p);
    S3D_LAUNCH_OK();
  }
  kq<<<grid, kFaBwdThreads, FaDqCfg<DH, 0>::kSmemBytes, stream>>>(t64, t64, t64, a.q, a.dout, p);
  S3D_LAUNCH_OK();
  return S3D_OK;
}

bool attn_tc_supported(int DH) { return DH == 48 || DH == 64 || DH == 96 || DH == 192; }

// Workspace of the single-score-pass backward: two bf16 [B*H, N, Npad] matrices. 0 = that path does not apply (head dims
// whose 64-column operand boxes would read a neighbouring head, short sequences where the two-kernel form is faster).
long long attn_bwd_workspace_bytes(int B, int H, int N, int DH) {
  const char* v = getenv("S3D_FA_SPILL_MIN_N");  // read per call: the parity tests lower it to cover this path at small N
  const int min_n = v == nullptr ? 128 : atoi(v);
  if (!attn_tc_fwd_supported(DH) || N < min_n || B <= 0 || H <= 0) return 0;
  const long long npad = (N + 63) / 64 * 64;
  const long long one = (2LL * B * H * N * npad + 1023) / 1024 * 1024;  // bytes of one matrix, 1 KiB aligned
  return 2 * one;
}
bool attn_tc_fwd_supported(int DH) { return attn_tc_supported(DH) || DH == 256; }

// head_dim 256 (timm Blocks of deit_base, N = 197): only the spill-ONLY form fits tensor memory; without a workspace the
// caller falls back to the mma.sync kernels
template <bool DROP>
static int fa_bwd_256(const AttnParams& a, cudaStream_t stream) {
  if (!fa_spill_applies(a, 256)) return S3D_ERR_UNSUPPORTED;
  FaBwdSetup st;
  if (int rc = fa_bwd_setup(a, 256, st, stream)) return rc;
  return fa_bwd_spill_launch<256, DROP, 2>(a, st, stream);
}

template <bool DROP>
static int fa_bwd_dispatch(const AttnParams& p, int DH, cudaStream_t stream) {
  switch (DH) {
    case 256: return fa_bwd_256<DROP>(p, stream);
    case 192: return fa_bwd_launch<192, DROP>(p, stream);
    case 96: return fa_bwd_launch<96, DROP>(p, stream);
    case 64: return fa_bwd_launch<64, DROP>(p, stream);
    case 48: return fa_bwd_launch<48, DROP>(p, stream);
    default: return S3D_ERR_UNSUPPORTED;
  }
}
// S3D_FA_FWD=2 selects the first-generation two-tile forward kernel (A operands in shared memory) for A/B measurements
template <bool DROP>
static int fa_fwd_dispatch(const AttnParams& p, int DH, cudaStream_t stream) {
  // With dropout the mask hashes sit on the single tile's softmax critical path (7.3 ms against 6.6 ms for the two-tile
  // kernel, whose second tile hides them, on the group_embed shape); without dropout the single-tile kernel wins
  // (5.6 against 6.1 ms). S3D_FA_FWD=1 / 2 forces one of them.
  static const int forced = []() { const char* v = getenv("S3D_FA_FWD"); return v == nullptr ? 0 : (v[0] == '2' ? 2 : (v[0] == '1' ? 1 : 0)); }();
  const bool two_tile = forced == 2 || (forced == 0 && DROP && p.N >= 1024);
  if (two_tile && DH != 256) {
    switch (DH) {
      case 192: return fa_fwd_launch<192, DROP>(p, stream);
      case 96: return fa_fwd_launch<96, DROP>(p, stream);
      case 64: return fa_fwd_launch<64, DROP>(p, stream);
      case 48: return fa_fwd_launch<48, DROP>(p, stream);
      default: return S3D_ERR_UNSUPPORTED;
    }
  }
  switch (DH) {
    case 256: return fa_fwd1_launch<256, DROP>(p, stream);
    case 192: return fa_fwd1_launch<192, DROP>(p, stream);
    case 96: return fa_fwd1_launch<96, DROP>(p, stream);
    case 64: return fa_fwd1_launch<64, DROP>(p, stream);
    case 48: return fa_fwd1_launch<48, DROP>(p, stream);
    default: return S3D_ERR_UNSUPPORTED;
  }
}

int attn_bwd_tc(const AttnParams& p, int DH, cudaStream_t stream) {
  if (p.B <= 0 || p.H <= 0 || p.N <= 0) return S3D_ERR_BAD_SHAPE;
  return p.drop_seed != nullptr ? fa_bwd_dispatch<true>(p, DH, stream) : fa_bwd_dispatch<false>(p, DH, stream);
}

int attn_fwd_tc(const AttnParams& p, int DH, cudaStream_t stream) {
  if (p.B <= 0 || p.H <= 0 || p.N <= 0) return S3D_ERR_BAD_SHAPE;
  return p.drop_seed != nullptr ? fa_fwd_dispatch<true>(p, DH, stream) : fa_fwd_dispatch<false>(p, DH, stream);
}

}  // namespace s3d
