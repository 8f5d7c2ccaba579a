// Scaled-dot-product attention core, forward and backward, for every sequence shape on the hot path:
//   timm-0.3.2 Attention.forward (q @ k^T * scale -> softmax -> @ v) with N in {15, 26, 197, 257, 513}, dh in {64, 256}
//   nn.MultiheadAttention inside the group_embed TransformerEncoderLayer (vit_3d_2d_pretrain.py:381,479): S = 12544,
//   dh = 192, sequence-first layout.
// Flash-style: scores never touch HBM (the reference materialises [B,H,N,N] fp32), online softmax in fp32, bf16 operands.
// Tensor-core path here is warp-level mma.sync.m16n8k16 with ldmatrix from XOR-swizzled shared memory; operands are
// addressed through (batch, head, row) strides so the timm [B,N,3,H,dh] and the sequence-first [S,Nb,3E] layouts are
// both consumed in place (no permute/contiguous copies).
#include "kernels.h"

#include <stdlib.h>

namespace s3d {

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// D = A * B with a zero accumulator (C is the zero register: no accumulator initialisation moves)
__device__ __forceinline__ void mma16816_z(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                           uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%10, %10, %10, %10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// smem tile of ROWS x DH bf16, 16-byte chunks XOR-swizzled by (row & 7): conflict-free ldmatrix in both orientations.
template <int DH>
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
  return base + (uint32_t)(((row * (DH / 8)) + (chunk ^ (row & 7))) << 4);
}

template <int DH, int ROWS>
__device__ __forceinline__ void load_tile(uint32_t sbase, const __nv_bfloat16* g, long long row_stride, int r0, int N,
                                          int tid, int nthr) {
  constexpr int CH = DH / 8;
  for (int i = tid; i < ROWS * CH; i += nthr) {
    const int r = i / CH, c = i % CH;
    const int gr = r0 + r;
    const int cr = gr < N ? gr : N - 1;
    cp_async16(tile_addr<DH>(sbase, r, c), g + (long long)cr * row_stride + c * 8, gr < N ? 16 : 0);
  }
}

// transpose of an 8x8 matrix of 16-bit elements spread over the warp (lane l holds row l/4, columns 2*(l%4)+{0,1})
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}

template <bool INDEP>
__device__ __forceinline__ void group_sync() {
  if (INDEP) __syncwarp(); else __syncthreads();
}

constexpr float kLog2e = 1.4426950408889634f;

// ================================================================================================
// Forward. CTA = NWARPS warps; each warp owns 16 query rows. INDEP: every warp works on its own (batch, head)
// (tiny sequences, N <= 16), otherwise the CTA shares K/V tiles of one (batch, head).
// ================================================================================================
template <int DH, int NWARPS, int BKV, bool INDEP>
__global__ void __launch_bounds__(NWARPS * 32) attn_fwd_kernel(const AttnParams p) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t smem_attn[];
  constexpr int BQ = INDEP ? 16 : 16 * NWARPS;
  constexpr int kQBytes = BQ * DH * 2;
  constexpr int kKVBytes = BKV * DH * 2;
  constexpr int NST = INDEP ? 1 : 2;                       // INDEP sequences fit one key tile: no ring needed
  constexpr int kPerGroup = kQBytes + NST * 2 * kKVBytes;  // Q + NST stages x (K, V)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tid = INDEP ? lane : threadIdx.x;
  constexpr int nthr = INDEP ? 32 : NWARPS * 32;
  const int wq = INDEP ? 0 : warp;

  long long bh;  // flattened (batch, head)
  int q_tile;
  if (INDEP) {
    bh = (long long)blockIdx.x * NWARPS + warp;
    q_tile = 0;
    if (bh >= (long long)p.B * p.H) return;  // whole warp exits; only __syncwarp is used below
  } else {
    bh = (long long)blockIdx.z * p.H + blockIdx.y;
    q_tile = blockIdx.x;
  }
  const int b = (int)(bh / p.H), h = (int)(bh % p.H);
  const uint32_t sbase = smem_u32(smem_attn) + (INDEP ? warp * kPerGroup : 0);
  const uint32_t sQ = sbase;
  const uint32_t sK0 = sbase + kQBytes;
  const uint32_t sV0 = sK0 + kKVBytes;

  const __nv_bfloat16* gq = p.q + (long long)b * p.qkv_bs + (long long)h * p.qkv_hs;
  const __nv_bfloat16* gk = p.k + (long long)b * p.qkv_bs + (long long)h * p.qkv_hs;
  const __nv_bfloat16* gv = p.v + (long long)b * p.qkv_bs + (long long)h * p.qkv_hs;
  const int q0 = q_tile * BQ;
  const int nkt = (p.N + BKV - 1) / BKV;

  load_tile<DH, BQ>(sQ, gq, p.qkv_rs, q0, p.N, tid, nthr);
  load_tile<DH, BKV>(sK0, gk, p.qkv_rs, 0, p.N, tid, nthr);
  load_tile<DH, BKV>(sV0, gv, p.qkv_rs, 0, p.N, tid, nthr);
  cp_async_commit();

  float o[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) { o[i][0] = 0.f; o[i][1] = 0.f; o[i][2] = 0.f; o[i][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const float sc = p.scale * kLog2e;

  for (int kt = 0; kt < nkt; ++kt) {
    const int st = INDEP ? 0 : (kt & 1);
    const uint32_t sK = sK0 + st * 2 * kKVBytes;
    const uint32_t sV = sK + kKVBytes;
    if (!INDEP && kt + 1 < nkt) {
      const uint32_t nK = sK0 + (st ^ 1) * 2 * kKVBytes;
      load_tile<DH, BKV>(nK, gk, p.qkv_rs, (kt + 1) * BKV, p.N, tid, nthr);
      load_tile<DH, BKV>(nK + kKVBytes, gv, p.qkv_rs, (kt + 1) * BKV, p.N, tid, nthr);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    group_sync<INDEP>();

    float s[BKV / 8][4];
#pragma unroll
    for (int i = 0; i < BKV / 8; ++i) { s[i][0] = 0.f; s[i][1] = 0.f; s[i][2] = 0.f; s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) {
      uint32_t a0, a1, a2, a3;
      ldsm_x4(tile_addr<DH>(sQ, wq * 16 + (lane & 15), ks * 2 + (lane >> 4)), a0, a1, a2, a3);
#pragma unroll
      for (int nt2 = 0; nt2 < BKV / 16; ++nt2) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(tile_addr<DH>(sK, nt2 * 16 + (lane & 7) + ((lane >> 4) << 3), ks * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
        mma16816(s[2 * nt2], a0, a1, a2, a3, b0, b1);
        mma16816(s[2 * nt2 + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    // mask keys beyond N, online softmax
    const int key0 = kt * BKV + 2 * (lane & 3);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < BKV / 8; ++nt) {
      const int kidx = key0 + nt * 8;
      if (kidx >= p.N) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
      if (kidx + 1 >= p.N) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float al0 = exp2f((m0 - mn0) * sc), al1 = exp2f((m1 - mn1) * sc);
    m0 = mn0; m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < BKV / 8; ++nt) {
      s[nt][0] = exp2f((s[nt][0] - mn0) * sc);
      s[nt][1] = exp2f((s[nt][1] - mn0) * sc);
      s[nt][2] = exp2f((s[nt][2] - mn1) * sc);
      s[nt][3] = exp2f((s[nt][3] - mn1) * sc);
      rs0 += s[nt][0] + s[nt][1];
      rs1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * al0 + rs0;
    l1 = l1 * al1 + rs1;
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) { o[i][0] *= al0; o[i][1] *= al0; o[i][2] *= al1; o[i][3] *= al1; }
    // O += P V
#pragma unroll
    for (int j = 0; j < BKV / 16; ++j) {
      const uint32_t a0 = pack_bf16x2(s[2 * j][0], s[2 * j][1]);
      const uint32_t a1 = pack_bf16x2(s[2 * j][2], s[2 * j][3]);
      const uint32_t a2 = pack_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1]);
      const uint32_t a3 = pack_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
      for (int dt2 = 0; dt2 < DH / 16; ++dt2) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(tile_addr<DH>(sV, j * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), dt2 * 2 + (lane >> 4)), b0, b1, b2, b3);
        mma16816(o[2 * dt2], a0, a1, a2, a3, b0, b1);
        mma16816(o[2 * dt2 + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    group_sync<INDEP>();  // all reads of this stage done before it is refilled
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float il0 = 1.f / l0, il1 = 1.f / l1;
  const int r0 = q0 + wq * 16 + (lane >> 2), r1 = r0 + 8;
  __nv_bfloat16* go = p.out + (long long)b * p.o_bs + (long long)h * p.o_hs;
#pragma unroll
  for (int nt = 0; nt < DH / 8; ++nt) {
    const int col = nt * 8 + 2 * (lane & 3);
    if (r0 < p.N) *reinterpret_cast<uint32_t*>(go + (long long)r0 * p.o_rs + col) = pack_bf16x2(o[nt][0] * il0, o[nt][1] * il0);
    if (r1 < p.N) *reinterpret_cast<uint32_t*>(go + (long long)r1 * p.o_rs + col) = pack_bf16x2(o[nt][2] * il1, o[nt][3] * il1);
  }
  if (p.lse != nullptr && (lane & 3) == 0) {
    float* gl = p.lse + bh * p.N;
    if (r0 < p.N) gl[r0] = m0 * p.scale + logf(l0);
    if (r1 < p.N) gl[r1] = m1 * p.scale + logf(l1);
  }
}

// ================================================================================================
// Backward, part 0: delta[b,h,i] = sum_d dO[i,d] * O[i,d]   (one warp per row)
// ================================================================================================
__global__ void __launch_bounds__(256) attn_delta_kernel(const AttnParams p, int DH) {
  pdl_prologue();
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)p.B * p.H * p.N;
  if (warp >= total) return;
  const int i = (int)(warp % p.N);
  const long long bh = warp / p.N;
  const int b = (int)(bh / p.H), h = (int)(bh % p.H);
  const long long off = (long long)b * p.o_bs + (long long)h * p.o_hs + (long long)i * p.o_rs;
  float s = 0.f;
  for (int d = lane * 2; d < DH; d += 64) {
    const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p.o + off + d));
    const float2 g = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p.dout + off + d));
    s += a.x * g.x + a.y * g.y;
  }
  s = warp_sum(s);
  if (lane == 0) p.delta[warp] = s;
}

// ================================================================================================
// Backward, part 1: dQ. Each warp owns 16 query rows and sweeps the key tiles, recomputing P from lse.
// ================================================================================================
template <int DH, int NWARPS, int BKV, bool INDEP>
__global__ void __launch_bounds__(NWARPS * 32) attn_bwd_dq_kernel(const AttnParams p) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t smem_attn[];
  constexpr int BQ = INDEP ? 16 : 16 * NWARPS;
  constexpr int kQBytes = BQ * DH * 2;
  constexpr int kKVBytes = BKV * DH * 2;
  constexpr int kPerGroup = 2 * kQBytes + 2 * kKVBytes;  // Q, dO, K, V (single stage)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tid = INDEP ? lane : threadIdx.x;
  constexpr int nthr = INDEP ? 32 : NWARPS * 32;
  const int wq = INDEP ? 0 : warp;
  long long bh;
  int q_tile;
  if (INDEP) {
    bh = (long long)blockIdx.x * NWARPS + warp;
    q_tile = 0;
    if (bh >= (long long)p.B * p.H) return;
  } else {
    bh = (long long)blockIdx.z * p.H + blockIdx.y;
    q_tile = blockIdx.x;
  }
  const int b = (int)(bh / p.H), h = (int)(bh % p.H);
  const uint32_t sbase = smem_u32(smem_attn) + (INDEP ? warp * kPerGroup : 0);
  const uint32_t sQ = sbase, sdO = sbase + kQBytes, sK = sbase + 2 * kQBytes, sV = sK + kKVBytes;
  const long long qoff = (long long)b * p.qkv_bs + (long long)h * p.qkv_hs;
  const long long ooff = (long long)b * p.o_bs + (long long)h * p.o_hs;
  const int q0 = q_tile * BQ;
  const int nkt = (p.N + BKV - 1) / BKV;

  load_tile<DH, BQ>(sQ, p.q + qoff, p.qkv_rs, q0, p.N, tid, nthr);
  load_tile<DH, BQ>(sdO, p.dout + ooff, p.o_rs, q0, p.N, tid, nthr);
  cp_async_commit();

  const int r0 = q0 + wq * 16 + (lane >> 2), r1 = r0 + 8;
  const float* glse = p.lse + bh * p.N;
  const float* gdel = p.delta + bh * p.N;
  // padded query rows: lse = +inf makes P = 0
  const float lse0 = r0 < p.N ? glse[r0] * kLog2e : INFINITY, lse1 = r1 < p.N ? glse[r1] * kLog2e : INFINITY;
  const float del0 = r0 < p.N ? gdel[r0] : 0.f, del1 = r1 < p.N ? gdel[r1] : 0.f;
  const float sc = p.scale * kLog2e;

  float dq[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) { dq[i][0] = 0.f; dq[i][1] = 0.f; dq[i][2] = 0.f; dq[i][3] = 0.f; }

  for (int kt = 0; kt < nkt; ++kt) {
    load_tile<DH, BKV>(sK, p.k + qoff, p.qkv_rs, kt * BKV, p.N, tid, nthr);
    load_tile<DH, BKV>(sV, p.v + qoff, p.qkv_rs, kt * BKV, p.N, tid, nthr);
    cp_async_commit();
    cp_async_wait<0>();
    group_sync<INDEP>();

    float s[BKV / 8][4], dp[BKV / 8][4];
#pragma unroll
    for (int i = 0; i < BKV / 8; ++i) {
      s[i][0] = 0.f; s[i][1] = 0.f; s[i][2] = 0.f; s[i][3] = 0.f;
      dp[i][0] = 0.f; dp[i][1] = 0.f; dp[i][2] = 0.f; dp[i][3] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) {
      uint32_t a0, a1, a2, a3, g0, g1, g2, g3;
      ldsm_x4(tile_addr<DH>(sQ, wq * 16 + (lane & 15), ks * 2 + (lane >> 4)), a0, a1, a2, a3);
      ldsm_x4(tile_addr<DH>(sdO, wq * 16 + (lane & 15), ks * 2 + (lane >> 4)), g0, g1, g2, g3);
#pragma unroll
      for (int nt2 = 0; nt2 < BKV / 16; ++nt2) {
        uint32_t b0, b1, b2, b3;
        const int krow = nt2 * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int kch = ks * 2 + ((lane >> 3) & 1);
        ldsm_x4(tile_addr<DH>(sK, krow, kch), b0, b1, b2, b3);
        mma16816(s[2 * nt2], a0, a1, a2, a3, b0, b1);
        mma16816(s[2 * nt2 + 1], a0, a1, a2, a3, b2, b3);
        ldsm_x4(tile_addr<DH>(sV, krow, kch), b0, b1, b2, b3);
        mma16816(dp[2 * nt2], g0, g1, g2, g3, b0, b1);
        mma16816(dp[2 * nt2 + 1], g0, g1, g2, g3, b2, b3);
      }
    }
    const int key0 = kt * BKV + 2 * (lane & 3);
#pragma unroll
    for (int nt = 0; nt < BKV / 8; ++nt) {
      const int kidx = key0 + nt * 8;
      const bool v0 = kidx < p.N, v1 = kidx + 1 < p.N;
      const float p00 = v0 ? exp2f(s[nt][0] * sc - lse0) : 0.f;
      const float p01 = v1 ? exp2f(s[nt][1] * sc - lse0) : 0.f;
      const float p10 = v0 ? exp2f(s[nt][2] * sc - lse1) : 0.f;
      const float p11 = v1 ? exp2f(s[nt][3] * sc - lse1) : 0.f;
      s[nt][0] = p00 * (dp[nt][0] - del0);
      s[nt][1] = p01 * (dp[nt][1] - del0);
      s[nt][2] = p10 * (dp[nt][2] - del1);
      s[nt][3] = p11 * (dp[nt][3] - del1);
    }
    // dQ += dS K
#pragma unroll
    for (int j = 0; j < BKV / 16; ++j) {
      const uint32_t a0 = pack_bf16x2(s[2 * j][0], s[2 * j][1]);
      const uint32_t a1 = pack_bf16x2(s[2 * j][2], s[2 * j][3]);
      const uint32_t a2 = pack_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1]);
      const uint32_t a3 = pack_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
      for (int dt2 = 0; dt2 < DH / 16; ++dt2) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(tile_addr<DH>(sK, j * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), dt2 * 2 + (lane >> 4)), b0, b1, b2, b3);
        mma16816(dq[2 * dt2], a0, a1, a2, a3, b0, b1);
        mma16816(dq[2 * dt2 + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    group_sync<INDEP>();
  }
  __nv_bfloat16* gdq = p.dq + qoff;
#pragma unroll
  for (int nt = 0; nt < DH / 8; ++nt) {
    const int col = nt * 8 + 2 * (lane & 3);
    if (r0 < p.N) *reinterpret_cast<uint32_t*>(gdq + (long long)r0 * p.qkv_rs + col) = pack_bf16x2(dq[nt][0] * p.scale, dq[nt][1] * p.scale);
    if (r1 < p.N) *reinterpret_cast<uint32_t*>(gdq + (long long)r1 * p.qkv_rs + col) = pack_bf16x2(dq[nt][2] * p.scale, dq[nt][3] * p.scale);
  }
}

// ================================================================================================
// Backward, part 2: dK, dV. Each warp owns 16 key rows and sweeps the query tiles with the transposed products
// S^T = K Q^T and dP^T = V dO^T so P^T / dS^T land directly in A-operand registers. SPLIT > 1 divides the dh output
// columns of the two accumulators across gridDim (register budget at dh = 192 / 256).
// ================================================================================================
template <int DH, int NWARPS, int BQ, int SPLIT, bool INDEP>
__global__ void __launch_bounds__(NWARPS * 32) attn_bwd_dkv_kernel(const AttnParams p) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t smem_attn[];
  constexpr int BKVT = INDEP ? 16 : 16 * NWARPS;
  constexpr int DHS = DH / SPLIT;
  constexpr int kKVBytes = BKVT * DH * 2;
  constexpr int kQBytes = BQ * DH * 2;
  constexpr int kPerGroup = 2 * kKVBytes + 2 * kQBytes + 2 * BQ * 4;  // K, V, Q, dO, lse, delta
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tid = INDEP ? lane : threadIdx.x;
  constexpr int nthr = INDEP ? 32 : NWARPS * 32;
  const int wk = INDEP ? 0 : warp;
  long long bh;
  int k_tile, split;
  if (INDEP) {
    const long long unit = (long long)blockIdx.x * NWARPS + warp;
    bh = unit / SPLIT;
    split = (int)(unit % SPLIT);
    k_tile = 0;
    if (bh >= (long long)p.B * p.H) return;
  } else {
    bh = (long long)blockIdx.z * p.H + blockIdx.y;
    k_tile = blockIdx.x / SPLIT;
    split = blockIdx.x % SPLIT;
  }
  const int b = (int)(bh / p.H), h = (int)(bh % p.H);
  uint8_t* sgen = smem_attn + (INDEP ? warp * kPerGroup : 0);
  const uint32_t sbase = smem_u32(sgen);
  const uint32_t sK = sbase, sV = sbase + kKVBytes, sQ = sV + kKVBytes, sdO = sQ + kQBytes;
  float* s_lse = reinterpret_cast<float*>(sgen + 2 * kKVBytes + 2 * kQBytes);
  float* s_del = s_lse + BQ;
  const long long qoff = (long long)b * p.qkv_bs + (long long)h * p.qkv_hs;
  const long long ooff = (long long)b * p.o_bs + (long long)h * p.o_hs;
  const int k0 = k_tile * BKVT;
  const int nqt = (p.N + BQ - 1) / BQ;
  const float* glse = p.lse + bh * p.N;
  const float* gdel = p.delta + bh * p.N;
  const float sc = p.scale * kLog2e;

  load_tile<DH, BKVT>(sK, p.k + qoff, p.qkv_rs, k0, p.N, tid, nthr);
  load_tile<DH, BKVT>(sV, p.v + qoff, p.qkv_rs, k0, p.N, tid, nthr);
  cp_async_commit();

  float dk[DHS / 8][4], dv[DHS / 8][4];
#pragma unroll
  for (int i = 0; i < DHS / 8; ++i) {
    dk[i][0] = 0.f; dk[i][1] = 0.f; dk[i][2] = 0.f; dk[i][3] = 0.f;
    dv[i][0] = 0.f; dv[i][1] = 0.f; dv[i][2] = 0.f; dv[i][3] = 0.f;
  }

  for (int qt = 0; qt < nqt; ++qt) {
    load_tile<DH, BQ>(sQ, p.q + qoff, p.qkv_rs, qt * BQ, p.N, tid, nthr);
    load_tile<DH, BQ>(sdO, p.dout + ooff, p.o_rs, qt * BQ, p.N, tid, nthr);
    cp_async_commit();
    for (int i = tid; i < BQ; i += nthr) {
      const int qi = qt * BQ + i;
      s_lse[i] = qi < p.N ? glse[qi] * kLog2e : INFINITY;
      s_del[i] = qi < p.N ? gdel[qi] : 0.f;
    }
    cp_async_wait<0>();
    group_sync<INDEP>();

    float st[BQ / 8][4], dpt[BQ / 8][4];
#pragma unroll
    for (int i = 0; i < BQ / 8; ++i) {
      st[i][0] = 0.f; st[i][1] = 0.f; st[i][2] = 0.f; st[i][3] = 0.f;
      dpt[i][0] = 0.f; dpt[i][1] = 0.f; dpt[i][2] = 0.f; dpt[i][3] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) {
      uint32_t a0, a1, a2, a3, g0, g1, g2, g3;
      ldsm_x4(tile_addr<DH>(sK, wk * 16 + (lane & 15), ks * 2 + (lane >> 4)), a0, a1, a2, a3);
      ldsm_x4(tile_addr<DH>(sV, wk * 16 + (lane & 15), ks * 2 + (lane >> 4)), g0, g1, g2, g3);
#pragma unroll
      for (int nt2 = 0; nt2 < BQ / 16; ++nt2) {
        uint32_t b0, b1, b2, b3;
        const int qrow = nt2 * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int qch = ks * 2 + ((lane >> 3) & 1);
        ldsm_x4(tile_addr<DH>(sQ, qrow, qch), b0, b1, b2, b3);
        mma16816(st[2 * nt2], a0, a1, a2, a3, b0, b1);
        mma16816(st[2 * nt2 + 1], a0, a1, a2, a3, b2, b3);
        ldsm_x4(tile_addr<DH>(sdO, qrow, qch), b0, b1, b2, b3);
        mma16816(dpt[2 * nt2], g0, g1, g2, g3, b0, b1);
        mma16816(dpt[2 * nt2 + 1], g0, g1, g2, g3, b2, b3);
      }
    }
    // P^T and dS^T (rows = keys, columns = queries). Padded queries have lse = +inf -> P = 0.
#pragma unroll
    for (int nt = 0; nt < BQ / 8; ++nt) {
      const int qc = nt * 8 + 2 * (lane & 3);
      const float l0 = s_lse[qc], l1 = s_lse[qc + 1];
      const float d0 = s_del[qc], d1 = s_del[qc + 1];
      const float p00 = exp2f(st[nt][0] * sc - l0);
      const float p01 = exp2f(st[nt][1] * sc - l1);
      const float p10 = exp2f(st[nt][2] * sc - l0);
      const float p11 = exp2f(st[nt][3] * sc - l1);
      st[nt][0] = p00; st[nt][1] = p01; st[nt][2] = p10; st[nt][3] = p11;
      dpt[nt][0] = p00 * (dpt[nt][0] - d0);
      dpt[nt][1] = p01 * (dpt[nt][1] - d1);
      dpt[nt][2] = p10 * (dpt[nt][2] - d0);
      dpt[nt][3] = p11 * (dpt[nt][3] - d1);
    }
    // dV += P^T dO ; dK += dS^T Q   (B operands: [k = queries][n = dh] row-major -> ldmatrix.trans)
#pragma unroll
    for (int j = 0; j < BQ / 16; ++j) {
      const uint32_t pa0 = pack_bf16x2(st[2 * j][0], st[2 * j][1]);
      const uint32_t pa1 = pack_bf16x2(st[2 * j][2], st[2 * j][3]);
      const uint32_t pa2 = pack_bf16x2(st[2 * j + 1][0], st[2 * j + 1][1]);
      const uint32_t pa3 = pack_bf16x2(st[2 * j + 1][2], st[2 * j + 1][3]);
      const uint32_t sa0 = pack_bf16x2(dpt[2 * j][0], dpt[2 * j][1]);
      const uint32_t sa1 = pack_bf16x2(dpt[2 * j][2], dpt[2 * j][3]);
      const uint32_t sa2 = pack_bf16x2(dpt[2 * j + 1][0], dpt[2 * j + 1][1]);
      const uint32_t sa3 = pack_bf16x2(dpt[2 * j + 1][2], dpt[2 * j + 1][3]);
      const int qrow = j * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
#pragma unroll
      for (int dt2 = 0; dt2 < DHS / 16; ++dt2) {
        uint32_t b0, b1, b2, b3;
        const int ch = split * (DHS / 8) + dt2 * 2 + (lane >> 4);
        ldsm_x4_t(tile_addr<DH>(sdO, qrow, ch), b0, b1, b2, b3);
        mma16816(dv[2 * dt2], pa0, pa1, pa2, pa3, b0, b1);
        mma16816(dv[2 * dt2 + 1], pa0, pa1, pa2, pa3, b2, b3);
        ldsm_x4_t(tile_addr<DH>(sQ, qrow, ch), b0, b1, b2, b3);
        mma16816(dk[2 * dt2], sa0, sa1, sa2, sa3, b0, b1);
        mma16816(dk[2 * dt2 + 1], sa0, sa1, sa2, sa3, b2, b3);
      }
    }
    group_sync<INDEP>();
  }
  const int r0 = k0 + wk * 16 + (lane >> 2), r1 = r0 + 8;
  __nv_bfloat16* gdk = p.dk + qoff;
  __nv_bfloat16* gdv = p.dv + qoff;
#pragma unroll
  for (int nt = 0; nt < DHS / 8; ++nt) {
    const int col = split * DHS + nt * 8 + 2 * (lane & 3);
    if (r0 < p.N) {
      *reinterpret_cast<uint32_t*>(gdk + (long long)r0 * p.qkv_rs + col) = pack_bf16x2(dk[nt][0] * p.scale, dk[nt][1] * p.scale);
      *reinterpret_cast<uint32_t*>(gdv + (long long)r0 * p.qkv_rs + col) = pack_bf16x2(dv[nt][0], dv[nt][1]);
    }
    if (r1 < p.N) {
      *reinterpret_cast<uint32_t*>(gdk + (long long)r1 * p.qkv_rs + col) = pack_bf16x2(dk[nt][2] * p.scale, dk[nt][3] * p.scale);
      *reinterpret_cast<uint32_t*>(gdv + (long long)r1 * p.qkv_rs + col) = pack_bf16x2(dv[nt][2], dv[nt][3]);
    }
  }
}

// ================================================================================================
// Backward for tiny sequences (N <= 16), fully fused: one warp owns one (batch, head) and produces delta, dQ, dK and dV
// from a single read of Q, K, V, dO (smem) and O (global). Replaces the delta + dQ + dK/dV kernel trio for the stage-1
// sequences of the group-embed model (12544 x 3 heads x 15 tokens x 12 layers): these are HBM-bound, so reading the
// operands once instead of twice is the whole win.
// ================================================================================================
// NST = 2: each warp double-buffers its four operand tiles (the next pair's loads fly while this one is multiplied).
// NST = 1: one set of tiles per warp and twice the warps per SM -- at head_dim 256 a set is 32 KB, so 3 double-buffered
// warps were all an SM could hold and their dependent ldmatrix -> MMA -> store chains left HBM at 62 % of the copy peak;
// six single-buffered warps keep as many bytes in flight and hide each other's chains.
template <int DH, int NWARPS, int NST>
__global__ void __launch_bounds__(NWARPS * 32) attn_bwd_small_kernel(const AttnParams p) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t smem_attn[];
  constexpr int kTile = 16 * DH * 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // persistent: warp w walks (batch, head) pairs w, w + W, ... with two smem stages, so the cp.async traffic of the
  // next pair is in flight while the current one is multiplied (the kernel is HBM-bound: 54 KB moved per pair)
  const long long BH = (long long)p.B * p.H;
  const long long gw = (long long)blockIdx.x * NWARPS + warp;
  const long long gstride = (long long)gridDim.x * NWARPS;
  const uint32_t wbase = smem_u32(smem_attn) + warp * NST * 4 * kTile;
  // one tile = 16 rows x DH bf16; lane l copies 16-byte chunk l (+32, ...) of every VALID row: the source pointer
  // advances by one row stride per row, the swizzled smem offsets (chunk ^ (row & 7)) are 8 per-lane constants.
  // Rows >= N are never written by cp.async; they are zeroed once here (P = 0 must not meet NaN garbage).
  constexpr int CH = DH / 8;
  uint32_t xs[(CH + 31) / 32][8];
#pragma unroll
  for (int c = 0; c < (CH + 31) / 32; ++c)
#pragma unroll
    for (int k = 0; k < 8; ++k) xs[c][k] = (uint32_t)(((c * 32 + lane) ^ k) << 4);
  for (int i = lane; i < NST * 4 * 16 * CH; i += 32) {
    const int t = i / (16 * CH), r = (i / CH) % 16, ch = i % CH;
    if (r >= p.N) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(tile_addr<DH>(wbase + t * kTile, r, ch)), "r"(0));
    }
  }
  __syncwarp();
  auto load16 = [&](uint32_t sb, const __nv_bfloat16* g, long long rs) {
#pragma unroll
    for (int c = 0; c < (CH + 31) / 32; ++c) {
      if (CH % 32 == 0 || c * 32 + lane < CH) {
        const __nv_bfloat16* src = g + (c * 32 + lane) * 8;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          if (r < p.N) {  // warp-uniform
            cp_async16(sb + r * (DH * 2) + xs[c][r & 7], src, 16);
            src += rs;
          }
        }
      }
    }
  };
  auto prefetch = [&](long long t, int st) {
    const int tb = (int)(t / p.H), th = (int)(t % p.H);
    const long long tq = (long long)tb * p.qkv_bs + (long long)th * p.qkv_hs;
    const long long to = (long long)tb * p.o_bs + (long long)th * p.o_hs;
    const uint32_t sb = wbase + st * 4 * kTile;
    load16(sb, p.q + tq, p.qkv_rs);
    load16(sb + kTile, p.k + tq, p.qkv_rs);
    load16(sb + 2 * kTile, p.v + tq, p.qkv_rs);
    load16(sb + 3 * kTile, p.dout + to, p.o_rs);
  };
  if (NST == 2) {
    if (gw < BH) prefetch(gw, 0);
    cp_async_commit();
  }
  int it = 0;
  for (long long bh = gw; bh < BH; bh += gstride, ++it) {
  if (NST == 2) {
    if (bh + gstride < BH) prefetch(bh + gstride, (it + 1) & 1);
  } else {
    prefetch(bh, 0);  // the previous pair's reads of these tiles ended at the __syncwarp closing its iteration
  }
  cp_async_commit();
  const int b = (int)(bh / p.H), h = (int)(bh % p.H);
  const uint32_t sbase = wbase + (NST == 2 ? (it & 1) : 0) * 4 * kTile;
  const uint32_t sQ = sbase, sK = sbase + kTile, sV = sbase + 2 * kTile, sdO = sbase + 3 * kTile;
  const long long qoff = (long long)b * p.qkv_bs + (long long)h * p.qkv_hs;

  // delta_i = sum_d dO[i,d] O[i,d] = sum_j P[i,j] dP[i,j] (O = P V, dP = dO V^T): taken from the P and dP fragments below
  // instead of re-reading O and dO from global memory row by row (that loop was 15 dependent load + warp-reduce round
  // trips per warp and dominated the kernel). Lane l owns rows r0 = l/4, r1 = r0+8 and key columns 2*(l%4)+{0,1} (+8).
  const float* glse = p.lse + bh * p.N;
  const int r0 = lane >> 2, r1 = r0 + 8;
  const int c0 = 2 * (lane & 3);
  // padded rows / columns: lse = +inf makes P = 0
  const float lse_r0 = r0 < p.N ? glse[r0] * kLog2e : INFINITY, lse_r1 = r1 < p.N ? glse[r1] * kLog2e : INFINITY;
  const float sc = p.scale * kLog2e;
  if (NST == 2) cp_async_wait<1>();
  else cp_async_wait<0>();
  __syncwarp();

  // ---- S = Q K^T, dP = dO V^T (rows = queries)
  float s[2][4] = {}, dp[2][4] = {};
#pragma unroll
  for (int ks = 0; ks < DH / 16; ++ks) {
    uint32_t a0, a1, a2, a3, g0, g1, g2, g3, b0, b1, b2, b3;
    ldsm_x4(tile_addr<DH>(sQ, lane & 15, ks * 2 + (lane >> 4)), a0, a1, a2, a3);
    ldsm_x4(tile_addr<DH>(sdO, lane & 15, ks * 2 + (lane >> 4)), g0, g1, g2, g3);
    const int krow = (lane & 7) + ((lane >> 4) << 3), kch = ks * 2 + ((lane >> 3) & 1);
    ldsm_x4(tile_addr<DH>(sK, krow, kch), b0, b1, b2, b3);
    mma16816(s[0], a0, a1, a2, a3, b0, b1);
    mma16816(s[1], a0, a1, a2, a3, b2, b3);
    ldsm_x4(tile_addr<DH>(sV, krow, kch), b0, b1, b2, b3);
    mma16816(dp[0], g0, g1, g2, g3, b0, b1);
    mma16816(dp[1], g0, g1, g2, g3, b2, b3);
  }
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const int kidx = nt * 8 + c0;
    const bool v0 = kidx < p.N, v1 = kidx + 1 < p.N;
    const float p00 = v0 ? exp2f(s[nt][0] * sc - lse_r0) : 0.f, p01 = v1 ? exp2f(s[nt][1] * sc - lse_r0) : 0.f;
    const float p10 = v0 ? exp2f(s[nt][2] * sc - lse_r1) : 0.f, p11 = v1 ? exp2f(s[nt][3] * sc - lse_r1) : 0.f;
    s[nt][0] = p00; s[nt][1] = p01; s[nt][2] = p10; s[nt][3] = p11;
  }
  // row sums over the 16 keys: 4 values per lane, then the 4 lanes of a quad
  float del_r0 = (s[0][0] * dp[0][0] + s[0][1] * dp[0][1]) + (s[1][0] * dp[1][0] + s[1][1] * dp[1][1]);
  float del_r1 = (s[0][2] * dp[0][2] + s[0][3] * dp[0][3]) + (s[1][2] * dp[1][2] + s[1][3] * dp[1][3]);
  del_r0 += __shfl_xor_sync(0xffffffffu, del_r0, 1);
  del_r1 += __shfl_xor_sync(0xffffffffu, del_r1, 1);
  del_r0 += __shfl_xor_sync(0xffffffffu, del_r0, 2);
  del_r1 += __shfl_xor_sync(0xffffffffu, del_r1, 2);
  // P as the A operand (rows = queries) of an m16n8k16 MMA; its transpose is taken below with movmatrix
  const uint32_t pq0 = pack_bf16x2(s[0][0], s[0][1]), pq1 = pack_bf16x2(s[0][2], s[0][3]);
  const uint32_t pq2 = pack_bf16x2(s[1][0], s[1][1]), pq3 = pack_bf16x2(s[1][2], s[1][3]);
  // dS = P * (dP - delta), pre-multiplied by the softmax scale that both dQ = scale * dS K and dK = scale * dS^T Q carry
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    s[nt][0] *= (dp[nt][0] - del_r0) * p.scale;
    s[nt][1] *= (dp[nt][1] - del_r0) * p.scale;
    s[nt][2] *= (dp[nt][2] - del_r1) * p.scale;
    s[nt][3] *= (dp[nt][3] - del_r1) * p.scale;
  }
  const bool ok0 = r0 < p.N, ok1 = r1 < p.N;
  const long long roff0 = qoff + (long long)r0 * p.qkv_rs + c0, roff1 = roff0 + 8 * p.qkv_rs;
  const uint32_t a0 = pack_bf16x2(s[0][0], s[0][1]), a1 = pack_bf16x2(s[0][2], s[0][3]);
  const uint32_t a2 = pack_bf16x2(s[1][0], s[1][1]), a3 = pack_bf16x2(s[1][2], s[1][3]);
  {  // ---- dQ = dS K
    uint32_t* q0 = reinterpret_cast<uint32_t*>(p.dq + roff0);
    uint32_t* q1 = reinterpret_cast<uint32_t*>(p.dq + roff1);
#pragma unroll
    for (int dt2 = 0; dt2 < DH / 16; ++dt2) {
      float acc0[4], acc1[4];
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(tile_addr<DH>(sK, (lane & 7) + (((lane >> 3) & 1) << 3), dt2 * 2 + (lane >> 4)), b0, b1, b2, b3);
      mma16816_z(acc0, a0, a1, a2, a3, b0, b1);
      mma16816_z(acc1, a0, a1, a2, a3, b2, b3);
      if (ok0) {
        q0[dt2 * 8] = pack_bf16x2(acc0[0], acc0[1]);
        q0[dt2 * 8 + 4] = pack_bf16x2(acc1[0], acc1[1]);
      }
      if (ok1) {
        q1[dt2 * 8] = pack_bf16x2(acc0[2], acc0[3]);
        q1[dt2 * 8 + 4] = pack_bf16x2(acc1[2], acc1[3]);
      }
    }
  }
  // ---- P^T and dS^T as A operands (rows = keys): a 16x16 fragment is four 8x8 blocks [[X00, X01], [X10, X11]] held
  // as (a0, a2 / a1, a3); its transpose is [[X00^T, X10^T], [X01^T, X11^T]], each block transposed across the warp by
  // movmatrix. This replaces recomputing K Q^T and V dO^T (64 of the kernel's 224 MMAs, 48 ldmatrix, 16 exp2).
  {  // ---- dV = P^T dO, dK = dS^T Q
    const uint32_t pa0 = movmatrix_trans(pq0), pa1 = movmatrix_trans(pq2), pa2 = movmatrix_trans(pq1),
                   pa3 = movmatrix_trans(pq3);
    const uint32_t sa0 = movmatrix_trans(a0), sa1 = movmatrix_trans(a2), sa2 = movmatrix_trans(a1),
                   sa3 = movmatrix_trans(a3);
    uint32_t* k0p = reinterpret_cast<uint32_t*>(p.dk + roff0);
    uint32_t* k1p = reinterpret_cast<uint32_t*>(p.dk + roff1);
    uint32_t* v0p = reinterpret_cast<uint32_t*>(p.dv + roff0);
    uint32_t* v1p = reinterpret_cast<uint32_t*>(p.dv + roff1);
    const int qrow = (lane & 7) + (((lane >> 3) & 1) << 3);
#pragma unroll
    for (int dt2 = 0; dt2 < DH / 16; ++dt2) {
      float v0[4], v1[4], k0[4], k1[4];
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(tile_addr<DH>(sdO, qrow, dt2 * 2 + (lane >> 4)), b0, b1, b2, b3);
      mma16816_z(v0, pa0, pa1, pa2, pa3, b0, b1);
      mma16816_z(v1, pa0, pa1, pa2, pa3, b2, b3);
      ldsm_x4_t(tile_addr<DH>(sQ, qrow, dt2 * 2 + (lane >> 4)), b0, b1, b2, b3);
      mma16816_z(k0, sa0, sa1, sa2, sa3, b0, b1);
      mma16816_z(k1, sa0, sa1, sa2, sa3, b2, b3);
      if (ok0) {
        v0p[dt2 * 8] = pack_bf16x2(v0[0], v0[1]);
        v0p[dt2 * 8 + 4] = pack_bf16x2(v1[0], v1[1]);
        k0p[dt2 * 8] = pack_bf16x2(k0[0], k0[1]);
        k0p[dt2 * 8 + 4] = pack_bf16x2(k1[0], k1[1]);
      }
      if (ok1) {
        v1p[dt2 * 8] = pack_bf16x2(v0[2], v0[3]);
        v1p[dt2 * 8 + 4] = pack_bf16x2(v1[2], v1[3]);
        k1p[dt2 * 8] = pack_bf16x2(k0[2], k0[3]);
        k1p[dt2 * 8 + 4] = pack_bf16x2(k1[2], k1[3]);
      }
    }
  }
  __syncwarp();  // all lanes are done with this stage before it is refilled two pairs later
  }
}

// ================================================================================================
// Forward for tiny sequences (N <= 16): one warp per (batch, head), persistent with two cp.async stages, same structure
// as attn_bwd_small_kernel. Replaces the generic kernel's INDEP path for the stage-1 sequences of the group-embed model
// (2820 warp instructions per (batch, head) there, 64 of them MMAs: index arithmetic and per-element rescaling).
// ================================================================================================
template <int DH, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) attn_fwd_small_kernel(const AttnParams p) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t smem_attn[];
  constexpr int kTile = 16 * DH * 2;
  constexpr int CH = DH / 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long BH = (long long)p.B * p.H;
  const long long gw = (long long)blockIdx.x * NWARPS + warp;
  const long long gstride = (long long)gridDim.x * NWARPS;
  const uint32_t wbase = smem_u32(smem_attn) + warp * 6 * kTile;  // two stages of {Q, K, V}
  uint32_t xs[(CH + 31) / 32][8];
#pragma unroll
  for (int c = 0; c < (CH + 31) / 32; ++c)
#pragma unroll
    for (int k = 0; k < 8; ++k) xs[c][k] = (uint32_t)(((c * 32 + lane) ^ k) << 4);
  for (int i = lane; i < 6 * 16 * CH; i += 32) {  // rows >= N are never written by cp.async: zero them once
    const int t = i / (16 * CH), r = (i / CH) % 16, ch = i % CH;
    if (r >= p.N) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(tile_addr<DH>(wbase + t * kTile, r, ch)), "r"(0));
    }
  }
  __syncwarp();
  auto load16 = [&](uint32_t sb, const __nv_bfloat16* g, long long rs) {
#pragma unroll
    for (int c = 0; c < (CH + 31) / 32; ++c) {
      if (CH % 32 == 0 || c * 32 + lane < CH) {
        const __nv_bfloat16* src = g + (c * 32 + lane) * 8;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          if (r < p.N) {  // warp-uniform
            cp_async16(sb + r * (DH * 2) + xs[c][r & 7], src, 16);
            src += rs;
          }
        }
      }
    }
  };
  auto prefetch = [&](long long t, int st) {
    const int tb = (int)(t / p.H), th = (int)(t % p.H);
    const long long tq = (long long)tb * p.qkv_bs + (long long)th * p.qkv_hs;
    const uint32_t sb = wbase + st * 3 * kTile;
    load16(sb, p.q + tq, p.qkv_rs);
    load16(sb + kTile, p.k + tq, p.qkv_rs);
    load16(sb + 2 * kTile, p.v + tq, p.qkv_rs);
  };
  if (gw < BH) prefetch(gw, 0);
  cp_async_commit();
  const int r0 = lane >> 2, r1 = r0 + 8;
  const int c0 = 2 * (lane & 3);
  const bool ok0 = r0 < p.N, ok1 = r1 < p.N;
  const float sc = p.scale * kLog2e;
  int it = 0;
  for (long long bh = gw; bh < BH; bh += gstride, ++it) {
    if (bh + gstride < BH) prefetch(bh + gstride, (it + 1) & 1);
    cp_async_commit();
    const int b = (int)(bh / p.H), h = (int)(bh % p.H);
    const uint32_t sQ = wbase + (it & 1) * 3 * kTile, sK = sQ + kTile, sV = sQ + 2 * kTile;
    cp_async_wait<1>();
    __syncwarp();
    // ---- S = Q K^T
    float s[2][4] = {};
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) {
      uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
      ldsm_x4(tile_addr<DH>(sQ, lane & 15, ks * 2 + (lane >> 4)), a0, a1, a2, a3);
      ldsm_x4(tile_addr<DH>(sK, (lane & 7) + ((lane >> 4) << 3), ks * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
      mma16816(s[0], a0, a1, a2, a3, b0, b1);
      mma16816(s[1], a0, a1, a2, a3, b2, b3);
    }
    // ---- softmax over the (<= 16) keys of each row: 4 values per lane and row, the 4 lanes of a quad share a row
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int kidx = nt * 8 + c0;
      s[nt][0] = kidx < p.N ? s[nt][0] * sc : -INFINITY;
      s[nt][1] = kidx + 1 < p.N ? s[nt][1] * sc : -INFINITY;
      s[nt][2] = kidx < p.N ? s[nt][2] * sc : -INFINITY;
      s[nt][3] = kidx + 1 < p.N ? s[nt][3] * sc : -INFINITY;
      m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
      m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - m0);
      s[nt][1] = exp2f(s[nt][1] - m0);
      s[nt][2] = exp2f(s[nt][2] - m1);
      s[nt][3] = exp2f(s[nt][3] - m1);
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    if (p.lse != nullptr && (lane & 3) == 0) {  // natural-log units of the scaled scores, as the backward kernels expect
      float* glse = p.lse + bh * p.N;
      if (ok0) glse[r0] = (m0 + log2f(l0)) * 0.6931471805599453f;
      if (ok1) glse[r1] = (m1 + log2f(l1)) * 0.6931471805599453f;
    }
    // ---- O = (P V) / l: the unnormalised exponentials (row maximum exactly 1) are the bf16 A operand and the division
    // happens in fp32 on the accumulator, as in the streaming kernels (one rounding fewer than normalising P first)
    const uint32_t a0 = pack_bf16x2(s[0][0], s[0][1]), a1 = pack_bf16x2(s[0][2], s[0][3]);
    const uint32_t a2 = pack_bf16x2(s[1][0], s[1][1]), a3 = pack_bf16x2(s[1][2], s[1][3]);
    const long long ooff = (long long)b * p.o_bs + (long long)h * p.o_hs + c0;
    uint32_t* o0 = reinterpret_cast<uint32_t*>(p.out + ooff + (long long)r0 * p.o_rs);
    uint32_t* o1 = reinterpret_cast<uint32_t*>(p.out + ooff + (long long)r1 * p.o_rs);
#pragma unroll
    for (int dt2 = 0; dt2 < DH / 16; ++dt2) {
      float acc0[4], acc1[4];
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(tile_addr<DH>(sV, (lane & 7) + (((lane >> 3) & 1) << 3), dt2 * 2 + (lane >> 4)), b0, b1, b2, b3);
      mma16816_z(acc0, a0, a1, a2, a3, b0, b1);
      mma16816_z(acc1, a0, a1, a2, a3, b2, b3);
      if (ok0) {
        o0[dt2 * 8] = pack_bf16x2(acc0[0] * i0, acc0[1] * i0);
        o0[dt2 * 8 + 4] = pack_bf16x2(acc1[0] * i0, acc1[1] * i0);
      }
      if (ok1) {
        o1[dt2 * 8] = pack_bf16x2(acc0[2] * i1, acc0[3] * i1);
        o1[dt2 * 8 + 4] = pack_bf16x2(acc1[2] * i1, acc1[3] * i1);
      }
    }
    __syncwarp();  // all lanes are done with this stage before it is refilled two pairs later
  }
}

// ------------------------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------------------------
static int check_attn(const AttnParams& p, int DH) {
  if (p.B <= 0 || p.H <= 0 || p.N <= 0) return S3D_ERR_BAD_SHAPE;
  if (DH != 64 && DH != 192 && DH != 256 && !attn_tc_supported(DH)) return S3D_ERR_UNSUPPORTED;
  if (p.qkv_rs % 8 || p.qkv_hs % 8 || p.qkv_bs % 8 || p.o_rs % 8 || p.o_hs % 8 || p.o_bs % 8) return S3D_ERR_ALIGNMENT;
  return S3D_OK;
}

template <typename K>
static int set_smem(K kern, int bytes) {
  S3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return S3D_OK;
}

template <int DH>
static int attn_fwd_dh(const AttnParams& p, cudaStream_t stream) {
  const long long BH = (long long)p.B * p.H;
  if (p.N <= 16) {  // persistent, one warp per (batch, head), two stages of {Q, K, V} per warp
    constexpr int NW = (DH == 64) ? 8 : 4;
    constexpr int smem = NW * 2 * 3 * 16 * DH * 2;
    auto kern = attn_fwd_small_kernel<DH, NW>;
    int rc = set_smem(kern, smem);
    if (rc) return rc;
    long long ctas = (BH + NW - 1) / NW;
    if (ctas > num_sms()) ctas = num_sms();
    S3D_CUDA_OK(launch_pdl(kern, dim3((unsigned)ctas), dim3(NW * 32), (size_t)(smem), stream, p));
  } else if (p.N <= 32) {
    constexpr int NW = 2, BKV = 32;
    constexpr int smem = 32 * DH * 2 + 4 * BKV * DH * 2;
    auto kern = attn_fwd_kernel<DH, NW, BKV, false>;
    int rc = set_smem(kern, smem);
    if (rc) return rc;
    if (p.B > 65535) return S3D_ERR_BAD_SHAPE;
    S3D_CUDA_OK(launch_pdl(kern, dim3(dim3(1, p.H, p.B)), dim3(NW * 32), (size_t)(smem), stream, p));
  } else {
    constexpr int NW = 4, BKV = (DH == 64) ? 64 : 32;
    constexpr int smem = 64 * DH * 2 + 4 * BKV * DH * 2;
    auto kern = attn_fwd_kernel<DH, NW, BKV, false>;
    int rc = set_smem(kern, smem);
    if (rc) return rc;
    if (p.B > 65535) return S3D_ERR_BAD_SHAPE;
    S3D_CUDA_OK(launch_pdl(kern, dim3(dim3((p.N + 63) / 64, p.H, p.B)), dim3(NW * 32), (size_t)(smem), stream, p));
  }
  S3D_LAUNCH_OK();
  return S3D_OK;
}

template <int DH>
static int attn_bwd_dh(const AttnParams& p, cudaStream_t stream) {
  const long long BH = (long long)p.B * p.H;
  if (p.N <= 16) {  // fused delta + dQ + dK + dV, one warp per (batch, head)
    static const bool two_stage = []() { const char* v = getenv("S3D_ATTN_SMALL_BWD_STAGES"); return v != nullptr && v[0] == '2'; }();
    if (DH == 256 && !two_stage) {  // single-buffered warps (see the kernel): measured 0.475 ms (3 x 2 stages) -> 0.407 ms
      // (6 warps, 5.7 TB/s) / 0.414 ms (7 warps) on the stage-1 shape of cfg3 (12544 x 3 heads x 15 tokens)
      static const bool six = []() { const char* v = getenv("S3D_ATTN_SMALL_BWD_WARPS"); return v == nullptr || v[0] != '7'; }();
      auto launch1 = [&](auto kern, int nw) -> int {
        const int smem = nw * 4 * 16 * DH * 2;
        int rc = set_smem(kern, smem);
        if (rc) return rc;
        long long ctas = (BH + nw - 1) / nw;
        if (ctas > num_sms()) ctas = num_sms();
        S3D_CUDA_OK(launch_pdl(kern, dim3((unsigned)ctas), dim3(nw * 32), (size_t)(smem), stream, p));
        S3D_LAUNCH_OK();
        return S3D_OK;
      };
      return six ? launch1(attn_bwd_small_kernel<DH, 6, 1>, 6) : launch1(attn_bwd_small_kernel<DH, 7, 1>, 7);
    }
    constexpr int NW = (DH == 64) ? 8 : (DH == 192 ? 4 : 3);
    constexpr int smem = NW * 2 * 4 * 16 * DH * 2;  // two stages of {Q, K, V, dO} per warp
    auto kern = attn_bwd_small_kernel<DH, NW, 2>;
    int rc = set_smem(kern, smem);
    if (rc) return rc;
    long long ctas = (BH + NW - 1) / NW;
    if (ctas > num_sms()) ctas = num_sms();
    S3D_CUDA_OK(launch_pdl(kern, dim3((unsigned)ctas), dim3(NW * 32), (size_t)(smem), stream, p));
    S3D_LAUNCH_OK();
    return S3D_OK;
  }
  {
    const long long rows = BH * p.N;
    S3D_CUDA_OK(launch_pdl(attn_delta_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), (size_t)(0), stream, p, DH));
    S3D_LAUNCH_OK();
  }
  constexpr int SPLIT = (DH == 64) ? 1 : 2;
  if (p.N <= 16) {
    constexpr int NW = (DH == 64) ? 4 : 2, BKV = 16;
    {
      constexpr int smem = NW * (2 * 16 * DH * 2 + 2 * BKV * DH * 2);
      auto kern = attn_bwd_dq_kernel<DH, NW, BKV, true>;
      int rc = set_smem(kern, smem);
      if (rc) return rc;
      S3D_CUDA_OK(launch_pdl(kern, dim3((unsigned)((BH + NW - 1) / NW)), dim3(NW * 32), (size_t)(smem), stream, p));
      S3D_LAUNCH_OK();
    }
    {
      constexpr int BQ = 16;
      constexpr int smem = NW * (2 * 16 * DH * 2 + 2 * BQ * DH * 2 + 2 * BQ * 4);
      auto kern = attn_bwd_dkv_kernel<DH, NW, BQ, SPLIT, true>;
      int rc = set_smem(kern, smem);
      if (rc) return rc;
      S3D_CUDA_OK(launch_pdl(kern, dim3((unsigned)((BH * SPLIT + NW - 1) / NW)), dim3(NW * 32), (size_t)(smem), stream, p));
      S3D_LAUNCH_OK();
    }
  } else if (p.N <= 32) {
    constexpr int NW = 2, BKV = 32;
    if (p.B > 65535) return S3D_ERR_BAD_SHAPE;
    {
      constexpr int smem = 2 * 32 * DH * 2 + 2 * BKV * DH * 2;
      auto kern = attn_bwd_dq_kernel<DH, NW, BKV, false>;
      int rc = set_smem(kern, smem);
      if (rc) return rc;
      S3D_CUDA_OK(launch_pdl(kern, dim3(dim3(1, p.H, p.B)), dim3(NW * 32), (size_t)(smem), stream, p));
      S3D_LAUNCH_OK();
    }
    {
      constexpr int BQ = 32;
      constexpr int smem = 2 * 32 * DH * 2 + 2 * BQ * DH * 2 + 2 * BQ * 4;
      auto kern = attn_bwd_dkv_kernel<DH, NW, BQ, SPLIT, false>;
      int rc = set_smem(kern, smem);
      if (rc) return rc;
      S3D_CUDA_OK(launch_pdl(kern, dim3(dim3(SPLIT, p.H, p.B)), dim3(NW * 32), (size_t)(smem), stream, p));
      S3D_LAUNCH_OK();
    }
  } else {
    constexpr int NW = 4;
    if (p.B > 65535) return S3D_ERR_BAD_SHAPE;
    {
      constexpr int BKV = (DH == 64) ? 64 : 32;
      constexpr int smem = 2 * 64 * DH * 2 + 2 * BKV * DH * 2;
      auto kern = attn_bwd_dq_kernel<DH, NW, BKV, false>;
      int rc = set_smem(kern, smem);
      if (rc) return rc;
      S3D_CUDA_OK(launch_pdl(kern, dim3(dim3((p.N + 63) / 64, p.H, p.B)), dim3(NW * 32), (size_t)(smem), stream, p));
      S3D_LAUNCH_OK();
    }
    {
      constexpr int BQ = (DH == 64) ? 64 : 32;
      constexpr int smem = 2 * 64 * DH * 2 + 2 * BQ * DH * 2 + 2 * BQ * 4;
      auto kern = attn_bwd_dkv_kernel<DH, NW, BQ, SPLIT, false>;
      int rc = set_smem(kern, smem);
      if (rc) return rc;
      S3D_CUDA_OK(launch_pdl(kern, dim3(dim3(((p.N + 63) / 64) * SPLIT, p.H, p.B)), dim3(NW * 32), (size_t)(smem), stream, p));
      S3D_LAUNCH_OK();
    }
  }
  return S3D_OK;
}

int attn_fwd(const AttnParams& p, int DH, cudaStream_t stream) {
  int rc = check_attn(p, DH);
  if (rc) return rc;
  if (p.q == nullptr || p.k == nullptr || p.v == nullptr || p.out == nullptr) return S3D_ERR_NULL;
  // long sequences run on the tcgen05 / TMEM flash kernel; everything else on the warp-level mma.sync kernels
  static const bool tc_enabled = []() { const char* v = getenv("S3D_ATTN_TC"); return v == nullptr || v[0] != '0'; }();
  // attention-probability dropout and the head dimensions 48 / 96 (group_embed of deit_tiny / deit_small) live in the
  // tcgen05 kernels only
  const bool tc_only = p.drop_seed != nullptr || (DH != 64 && DH != 192 && DH != 256);
  if (tc_only) return attn_tc_supported(DH) ? attn_fwd_tc(p, DH, stream) : S3D_ERR_UNSUPPORTED;
  // every sequence of at least one 64-key block runs on the tcgen05 kernels (timm Block shapes N = 197 / 257 / 513 and the
  // group_embed layer); the 15- and 26-token sequences keep the persistent warp-per-sequence mma.sync kernels
  if (tc_enabled && attn_tc_fwd_supported(DH) && p.N >= 64) {
    const int rc_tc = attn_fwd_tc(p, DH, stream);
    if (rc_tc != S3D_ERR_UNSUPPORTED) return rc_tc;
  }
  switch (DH) {
    case 64: return attn_fwd_dh<64>(p, stream);
    case 192: return attn_fwd_dh<192>(p, stream);
    default: return attn_fwd_dh<256>(p, stream);
  }
}

int attn_bwd(const AttnParams& p, int DH, cudaStream_t stream) {
  int rc = check_attn(p, DH);
  if (rc) return rc;
  if (p.q == nullptr || p.k == nullptr || p.v == nullptr || p.o == nullptr || p.dout == nullptr || p.dq == nullptr ||
      p.dk == nullptr || p.dv == nullptr || p.lse == nullptr || p.delta == nullptr)
    return S3D_ERR_NULL;
  static const bool tc_enabled = []() { const char* v = getenv("S3D_ATTN_TC"); return v == nullptr || v[0] != '0'; }();
  const bool tc_only = p.drop_seed != nullptr || (DH != 64 && DH != 192 && DH != 256);
  const bool spill = p.workspace != nullptr && attn_bwd_workspace_bytes(p.B, p.H, p.N, DH) > 0 &&
                     p.workspace_bytes >= attn_bwd_workspace_bytes(p.B, p.H, p.N, DH);
  // head_dim 256 runs on tcgen05 only in the spill-only form (dQ as a third GEMM), i.e. only with a workspace
  const bool tc_ok = attn_tc_supported(DH) || (DH == 256 && spill);
  if (tc_only) return tc_ok ? attn_bwd_tc(p, DH, stream) : S3D_ERR_UNSUPPORTED;
  // Backward on tcgen05: with a workspace (single score pass: one flash kernel that also spills P o mask / dS, then two
  // batched GEMMs) from N = 128 -- measured in the cfg4 / cfg5 steps (N = 257 / 513, dh 64): 1.82 / 1.37 ms per step
  // against 2.12 / 1.62 ms for the mma.sync kernels. Without one it is three kernels (dK, dV, dQ) whose per-CTA prologue
  // (operand tile -> tensor memory, ring start-up) only amortises over long sequences (2.61 / 1.84 ms on the same
  // shapes): from N = 1024. head_dim 256 does not fit the TMEM budget of either form.
  static const int tc_min_n = []() { const char* v = getenv("S3D_ATTN_TC_BWD_MIN_N"); return v == nullptr ? 1024 : atoi(v); }();
  if (tc_enabled && tc_ok && (spill || (DH != 256 && p.N >= tc_min_n))) {
    const int rc_tc = attn_bwd_tc(p, DH, stream);
    if (rc_tc != S3D_ERR_UNSUPPORTED) return rc_tc;
  }
  switch (DH) {
    case 64: return attn_bwd_dh<64>(p, stream);
    case 192: return attn_bwd_dh<192>(p, stream);
    default: return attn_bwd_dh<256>(p, stream);
  }
}

}  // namespace s3d
