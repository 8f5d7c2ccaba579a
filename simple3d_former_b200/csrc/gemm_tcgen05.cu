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
//
// L2 -> SM operand traffic is what bounds a 128 x BN tile (48 KB per 4.2 MFLOP at BN = 256 hits the ~6.9 TB/s L2 cap at
// ~580 TFLOP/s), so CTAs are launched as clusters of CL along M that share the B tile: every CTA loads 1/CL of B and
// TMA-multicasts it to all CL shared memories; smem slots are released with a multicast tcgen05.commit so a producer
// only overwrites a slot once every CTA of the cluster has consumed it.
// Split-K (weight-gradient GEMMs: few output tiles, K = tokens) gives each work item a K range and reduces with
// red.global.add.v4.f32 into the fp32 output.
#include "kernels.h"

#include <stdlib.h>
#include <type_traits>
#include <string.h>

namespace s3d {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarp0 = 4;

// EPW = epilogue warps per TMEM lane quadrant: 2 (8 warps) by default, 4 (16 warps) for the activation-GRADIENT
// epilogues (dGELU, dReLU: global aux read -> ~10 packed instructions per element -> pack -> staging -> TMA store is a
// dependent chain; four warps per scheduler hide more of it). Their 64 KB of staging tiles cost one operand stage at
// BN = 256, which is why the other epilogues keep 8 warps.
template <int BN, int EPW>
struct GemmCfg {
  static_assert(EPW == 2 || (EPW == 4 && BN == 256), "a part of the tile must span whole 128-byte staging rows");
  static constexpr int kEpiWarps = 4 * EPW;
  static constexpr int kThreads = 32 * (kEpiWarp0 + kEpiWarps);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int kStagingBytes = (BN >= 128) ? kEpiWarps * 4096 : 0;  // TMA-store staging (BN = 64 stores directly)
  static constexpr int kMaxStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kFit = (232448 - kStagingBytes - 1024 - 256) / kStageBytes;
  static constexpr int kStages = kFit < kMaxStages ? kFit : kMaxStages;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align*/ + 256 /*barriers*/;
};


// ---------------------------------------- epilogue helpers ----------------------------------------
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// accumulator chunk (32 columns of this thread's row) -> alpha, bias
__device__ __forceinline__ void epi_load(const GemmParams& p, uint32_t taddr, int n, int split, float (&f)[32]) {
  uint32_t v[32];
  tmem_ld_32x32b_x32(taddr, v);
  tc_wait_ld();
  const bool full = (n + 32 <= p.N);
  if (p.bias != nullptr && split == 0 && full) {  // alpha * acc + bias on packed pairs
    const float2 al = bcast2(p.alpha);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
      const float2 lo = ffma2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), al, make_float2(b4.x, b4.y));
      const float2 hi = ffma2(make_float2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), al, make_float2(b4.z, b4.w));
      f[j] = lo.x; f[j + 1] = lo.y; f[j + 2] = hi.x; f[j + 3] = hi.y;
    }
    return;
  }
  if (p.alpha == 1.0f) {  // warp-uniform: the activation-gradient GEMMs carry no scale
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * p.alpha;
  }
  if (p.bias != nullptr && split == 0 && n < p.N) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
        f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
      }
    } else {
      for (int j = 0; j < 32 && n + j < p.N; ++j) f[j] += __ldg(p.bias + n + j);
    }
  }
}

// activation / activation-gradient part of the epilogue (GELU's pre-activation save is handled by the caller)
__device__ __forceinline__ void epi_act(const GemmParams& p, int row, bool row_ok, int n, float (&f)[32]) {
  const bool full = (n + 32 <= p.N);
  if (p.epilogue == EPI_GELU) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float2 g2 = gelu_erf2(make_float2(f[j], f[j + 1]));
      f[j] = g2.x;
      f[j + 1] = g2.y;
    }
  } else if (p.epilogue == EPI_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
  } else if ((p.epilogue == EPI_DGELU || p.epilogue == EPI_DRELU) && row_ok && n < p.N) {
    const __nv_bfloat16* ai = p.aux_in + (long long)row * p.ld_aux_in + n;
    float a[32];
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        const uint4 u = *reinterpret_cast<const uint4*>(ai + j);
        const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
        a[j] = a0.x; a[j + 1] = a0.y; a[j + 2] = a1.x; a[j + 3] = a1.y;
        a[j + 4] = a2.x; a[j + 5] = a2.y; a[j + 6] = a3.x; a[j + 7] = a3.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) a[j] = (n + j < p.N) ? __bfloat162float(ai[j]) : 0.f;
    }
    if (p.epilogue == EPI_DGELU) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float2 g2 = fmul2(make_float2(f[j], f[j + 1]), gelu_erf_grad2(make_float2(a[j], a[j + 1])));
        f[j] = g2.x;
        f[j + 1] = g2.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = a[j] > 0.f ? f[j] : 0.f;
    }
  }
}

// Epilogue operands (bf16 aux of dGELU / dReLU, fp32 residual) through the warp's staging tile: the natural layout of
// the epilogue is lane = row, so a direct read makes every 16-byte load of a warp touch 32 different 128-byte lines
// (~32 LSU cycles per instruction, ~4 k cycles per 128 x 256 tile -- the dGELU GEMM ran at 48 % tensor pipe). Here the
// warp loads its 32 x 128-byte operand tile with COALESCED loads (8 lanes per row), parks it in the swizzled staging tile
// and every lane then reads its own row from shared memory.
__device__ __forceinline__ void epi_operand_to_stage(uint8_t* stage, const uint8_t* src, long long pitch_bytes, int rows_ok,
                                                     int lane) {
  uint4 t[8];
  const int piece = lane & 7, r0 = lane >> 3;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + r0;
    t[i] = rr < rows_ok ? __ldg(reinterpret_cast<const uint4*>(src + rr * pitch_bytes + piece * 16)) : make_uint4(0u, 0u, 0u, 0u);
  }
  if (lane == 0) tma_store_wait_read();  // the previous TMA store has finished reading the tile
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + r0;
    *reinterpret_cast<uint4*>(stage + rr * 128 + ((piece ^ (rr & 7)) << 4)) = t[i];
  }
  __syncwarp();
}
__device__ __forceinline__ void epi_act_grad_smem(const GemmParams& p, const uint8_t* srow, int r, int hh, float (&f)[32]) {
  if (p.epilogue == EPI_DGELU) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 u = *reinterpret_cast<const uint4*>(srow + (((hh * 4 + q) ^ (r & 7)) << 4));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = q * 8 + 2 * i;
        const float2 g2 = fmul2(make_float2(f[j], f[j + 1]), gelu_erf_grad2_clamped(unpack_bf16x2(clamp_bf16x2_4p5(w[i]))));
        f[j] = g2.x;
        f[j + 1] = g2.y;
      }
    }
    return;
  }
  float a[32];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 u = *reinterpret_cast<const uint4*>(srow + (((hh * 4 + q) ^ (r & 7)) << 4));
    const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
    a[q * 8] = a0.x; a[q * 8 + 1] = a0.y; a[q * 8 + 2] = a1.x; a[q * 8 + 3] = a1.y;
    a[q * 8 + 4] = a2.x; a[q * 8 + 5] = a2.y; a[q * 8 + 6] = a3.x; a[q * 8 + 7] = a3.y;
  }
  {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = a[j] > 0.f ? f[j] : 0.f;
  }
}
__device__ __forceinline__ void epi_residual_smem(const uint8_t* srow, int r, float (&f)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 rv = *reinterpret_cast<const float4*>(srow + ((q ^ (r & 7)) << 4));
    f[q * 4] += rv.x; f[q * 4 + 1] += rv.y; f[q * 4 + 2] += rv.z; f[q * 4 + 3] += rv.w;
  }
}

__device__ __forceinline__ void prefetch_l2(const void* ptr) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}

__device__ __forceinline__ void epi_residual(const GemmParams& p, long long boff_r, int row, bool row_ok, int n,
                                             float (&f)[32]) {
  if (p.residual == nullptr || !row_ok || n >= p.N) return;
  const float* r = p.residual + boff_r + (long long)row * p.ldr + n;
  if (n + 32 <= p.N) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 rv = *reinterpret_cast<const float4*>(r + j);
      f[j] += rv.x; f[j + 1] += rv.y; f[j + 2] += rv.z; f[j + 3] += rv.w;
    }
  } else {
    for (int j = 0; j < 32 && n + j < p.N; ++j) f[j] += r[j];
  }
}

// staging tile of one epilogue warp: 32 rows x 128 bytes, 16-byte chunks XOR-swizzled by (row & 7) (= TMA SWIZZLE_128B)
__device__ __forceinline__ void stage_bf16(uint8_t* srow, int r, int half, const float (&f)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]);
    u.y = pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]);
    u.z = pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]);
    u.w = pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]);
    *reinterpret_cast<uint4*>(srow + (((half * 4 + q) ^ (r & 7)) << 4)) = u;
  }
}
__device__ __forceinline__ void stage_f32(uint8_t* srow, int r, const float (&f)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(srow + ((q ^ (r & 7)) << 4)) = make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
}

template <int BN, int A_MN, int B_MN, int CL, int EPW>
__global__ void __launch_bounds__(GemmCfg<BN, EPW>::kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_d, const __grid_constant__ CUtensorMap tma_aux,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN, EPW>;
  pdl_trigger();  // the next grid may be scheduled (and run its on-chip prologue) while this one works
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem + Cfg::kStages * Cfg::kStageBytes;  // one 4 KB TMA-store staging tile per epilogue warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  static_assert(CL == 1 || (BN / 64) % CL == 0, "every CTA of the cluster loads BN / CL rows (whole 64-wide chunks)");
  const int num_m = (p.M + BM - 1) / BM;
  const int num_msuper = (num_m + CL - 1) / CL;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_kb_total = (p.K + BK - 1) / BK;
  const int splits = p.splits;
  const int kb_per_split = (num_kb_total + splits - 1) / splits;
  const int items_per_batch = num_msuper * num_n * splits;
  const int num_items = items_per_batch * (p.batched ? p.batch : 1);  // batches are folded into the persistent item loop
  const int crank = (CL > 1) ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CL;
  const int num_clusters = gridDim.x / CL;
  constexpr uint16_t kMcMask = (uint16_t)((1u << CL) - 1);
  constexpr int kBRowsPerCta = BN / CL;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CL);  // one multicast commit from every CTA of the cluster
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], Cfg::kEpiWarps);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // peers' barriers are initialised before any multicast can target them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above touched only on-chip state and kernel parameters and may overlap the previous grid's tail;
  // from here on global memory is read / written
  pdl_wait();

  // work item -> (m tile of this CTA, n tile, k-block range)
  // batch index -> tensor-map batch coordinate of an operand: (bt / inner) * mul + bt % inner (a two-level (batch, head)
  // offset whose strides are both multiples of the head stride collapses into ONE coordinate of stride head_stride).
  // Integer divisions cost ~40 issued instructions each and every role decodes every item: the common cases (no batch,
  // no split-K) take none, the tile split takes one.
  auto bcoord = [&](int bt, int mul) {
    if (p.batch_inner == 1) return bt * mul;
    const int hi = bt / p.batch_inner;
    return hi * mul + (bt - hi * p.batch_inner);
  };
  auto decode = [&](int item, int& m0, int& n0, int& kb0, int& kb1, int& split, int& bt) {
    bt = 0;
    if (p.batched && p.batch > 1) {
      bt = item / items_per_batch;
      item -= bt * items_per_batch;
    }
    int t = item;
    split = 0;
    if (splits > 1) {
      t = item / splits;
      split = item - t * splits;
    }
    const int tq = t / num_n;
    n0 = (t - tq * num_n) * BN;
    m0 = (tq * CL + crank) * BM;
    kb0 = split * kb_per_split;
    kb1 = min(num_kb_total, kb0 + kb_per_split);
  };

  if (warp == 0) {
    // ------------------------------------------ TMA producer ------------------------------------------
    // One thread feeds the ring. Its loop must stay lean: a weight-gradient GEMM (both operands MN-major) issues four
    // 8 KB boxes per k-block and ran 10 % slower when every load picked its tensor-map rank at run time (0.675 against
    // 0.613 ms on M = 3072, N = 768, K = 188160). All operand maps are rank 3 (batch extent 1 when there is none); the
    // panel layout of the attention workspace (rank 4) is a compile-time variant of the whole loop.
    auto produce = [&](auto a_tag, auto b_tag, auto pk_tag) {
      // MN-major operands through a rank-4 "chunk view" (make_tmap_bf16_chunks / _panel): ONE load brings all the
      // 64-wide chunks of the tile instead of one 8 KB box per chunk
      constexpr bool kChunkA = decltype(a_tag)::value, kChunkB = decltype(b_tag)::value;
      constexpr bool kPanelK = decltype(pk_tag)::value;  // K-major A read from 64-column panels (panel index = k-block)
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        int m0, n0, kb0, kb1, split, bt;
        decode(item, m0, n0, kb0, kb1, split, bt);
        const int batch = bcoord(bt, p.bmul_a), batch_b = bcoord(bt, p.bmul_b);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          const int k0 = kb * BK;
          // A: this CTA's own 128 rows
          if (A_MN) {
            if (kChunkA) {
              tma_load_4d(sa, &tma_a, &full_bar[stage], 0, k0, m0 >> 6, batch);
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i) tma_load_3d(sa + i * 8192, &tma_a, &full_bar[stage], m0 + 64 * i, k0, batch);
            }
          } else if (kPanelK) {
            tma_load_4d(sa, &tma_a, &full_bar[stage], 0, m0, kb, batch);
          } else {
            tma_load_3d(sa, &tma_a, &full_bar[stage], k0, m0, batch);
          }
          // B: 1/CL of the tile, multicast to every CTA of the cluster
          if (B_MN) {
            constexpr int kCh = BN / 64 / CL;  // chunks this CTA loads
            if (kChunkB) {
              uint8_t* dst = sb + crank * kCh * 8192;
              if (CL > 1) tma_load_4d_mc(dst, &tma_b, &full_bar[stage], 0, k0, (n0 >> 6) + crank * kCh, batch_b, kMcMask);
              else tma_load_4d(dst, &tma_b, &full_bar[stage], 0, k0, (n0 >> 6) + crank * kCh, batch_b);
            } else {
#pragma unroll
              for (int i = 0; i < kCh; ++i) {
                const int ch = crank * kCh + i;
                if (CL > 1) tma_load_3d_mc(sb + ch * 8192, &tma_b, &full_bar[stage], n0 + 64 * ch, k0, batch_b, kMcMask);
                else tma_load_3d(sb + ch * 8192, &tma_b, &full_bar[stage], n0 + 64 * ch, k0, batch_b);
              }
            }
          } else {
            uint8_t* dst = sb + crank * kBRowsPerCta * 128;
            const int nrow = n0 + crank * kBRowsPerCta;
            if (CL > 1) tma_load_3d_mc(dst, &tma_b, &full_bar[stage], k0, nrow, batch_b, kMcMask);
            else tma_load_3d(dst, &tma_b, &full_bar[stage], k0, nrow, batch_b);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    };
    if (lane == 0) {
      // chunk views need whole 64-wide chunks (M / N multiples of 64); both flags are set together by the host
      using T = std::true_type;
      using F = std::false_type;
      if (!A_MN && p.a_panel == 2) {
        if (B_MN && p.chunk_b) produce(F{}, T{}, T{});
        else produce(F{}, F{}, T{});
      } else if (A_MN && B_MN && p.chunk_a && p.chunk_b) produce(T{}, T{}, F{});
      else if (A_MN && p.chunk_a) produce(T{}, F{}, F{});
      else if (B_MN && p.chunk_b) produce(F{}, T{}, F{});
      else produce(F{}, F{}, F{});
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------- MMA issuer -------------------------------------------
    // warp-converged control flow, one elected lane issues; descriptor low words precomputed per stage
    {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      const uint32_t hi_a = smem_desc_hi_sw128(A_MN ? p.mn_sbo : p.k_sbo), hi_b = smem_desc_hi_sw128(B_MN ? p.mn_sbo : p.k_sbo);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem), A_MN ? p.mn_lbo : p.k_lbo);
      const uint32_t b_lo0 = smem_desc_lo(smem_u32(smem) + Cfg::kABytes, B_MN ? p.mn_lbo : p.k_lbo);
      constexpr uint32_t kStepA = (A_MN ? 2048 : 32) >> 4, kStepB = (B_MN ? 2048 : 32) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        int m0, n0, kb0, kb1, split, bt;
        decode(item, m0, n0, kb0, kb1, split, bt);
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (Cfg::kStageBytes >> 4), b_lo = b_lo0 + stage * (Cfg::kStageBytes >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_f16_ss2(d_tmem, a_lo + k * kStepA, hi_a, b_lo + k * kStepB, hi_b, idesc, (kb > kb0) || (k != 0));
            // frees the smem slot (in every CTA of the cluster) once these MMAs retire
            if (CL > 1) umma_commit_mc(&empty_bar[stage], kMcMask);
            else umma_commit(&empty_bar[stage]);
            if (kb + 1 == kb1) umma_commit(&tmem_full[as]);  // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (kb1 <= kb0) {  // empty split range: still hand the (unused) accumulator stage to the epilogue
          if (elect_one()) umma_commit(&tmem_full[as]);
          __syncwarp();
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // -------------------------------------------- epilogue --------------------------------------------
    const int e = warp - kEpiWarp0;  // 0 .. 4 * EPW - 1
    const int quad = warp & 3;       // TMEM lane quadrant this warp may access
    const int half = e >> 2;         // which part (1 / EPW) of the BN columns
    constexpr int kColsPerHalf = BN / EPW;
    int as = 0;
    uint32_t aphase = 0;
    for (int item = cluster_id; item < num_items; item += num_clusters) {
      int m0, n0, kb0, kb1, split, bt;
      decode(item, m0, n0, kb0, kb1, split, bt);
      const int batch = bcoord(bt, p.bmul_d);
      const long long boff_d = p.batched ? (long long)batch * p.batch_stride_d : 0;
      const long long boff_r = p.batched ? (long long)batch * p.batch_stride_r : 0;
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < p.M;
      if (splits == 1 && row_ok) {
        // The epilogue reads its residual / aux operands straight from global memory, one 32-column chunk at a time,
        // and each read used to expose a full DRAM round trip (short-K GEMMs were epilogue-latency bound: the fp32
        // residual proj GEMM ran at 565 TFLOP/s, the dGELU GEMM at 790). Pull this thread's lines of the tile into L2
        // now, while the tensor cores are still producing the accumulator.
        const int nb0 = n0 + half * kColsPerHalf;
        if (p.residual != nullptr) {
          const float* r = p.residual + boff_r + (long long)row * p.ldr + nb0;
#pragma unroll
          for (int c = 0; c < kColsPerHalf; c += 32)
            if (nb0 + c < p.N) prefetch_l2(r + c);
        }
        if (p.aux_in != nullptr) {
          const __nv_bfloat16* a = p.aux_in + (long long)row * p.ld_aux_in + nb0;
#pragma unroll
          for (int c = 0; c < kColsPerHalf; c += 64)
            if (nb0 + c < p.N) prefetch_l2(a + c);
        }
      }
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      if constexpr (BN >= 128) {
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN);
        if (splits == 1) {
          // ---- staged epilogue: registers -> swizzled smem -> TMA store (full 128-byte lines, hardware tail clipping).
          // One 32-column chunk is live at a time (small code: short launches are instruction-fetch sensitive).
          uint8_t* stage = stage_base + e * 4096;
          uint8_t* srow = stage + lane * 128;
          const int nh = p.out_fp32 ? 1 : 2;  // 32-column chunks per 128-byte staging row
          const int m_box = m0 + quad * 32;
#pragma unroll 1
          for (int c = 0; c < kColsPerHalf; c += 32 * nh) {
            const int col0 = half * kColsPerHalf + c;
            const int n = n0 + col0;
            if (n >= p.N) break;  // warp-uniform
            if (p.epilogue == EPI_GELU && p.aux_out != nullptr && !p.out_fp32) {
              // fc1 of the timm Mlp: the pre-activation (bf16, saved for the backward pass) and GELU(pre) leave through the
              // same staging tile. The accumulator chunk is read from tensor memory ONCE: the pre-activation is staged
              // at once, the activation waits as 32 packed registers until the first store has read the tile.
              if (lane == 0) tma_store_wait_read();
              __syncwarp();
              uint32_t act0[16], act1[16];
              {
                float f[32];
                epi_load(p, trow + col0, n, split, f);
                stage_bf16(srow, lane, 0, f);
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  const float2 g2 = gelu_erf2(make_float2(f[j], f[j + 1]));
                  act0[j >> 1] = pack_bf16x2(g2.x, g2.y);
                }
              }
              {
                float f[32];
                epi_load(p, trow + col0 + 32, n + 32, split, f);
                stage_bf16(srow, lane, 1, f);
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  const float2 g2 = gelu_erf2(make_float2(f[j], f[j + 1]));
                  act1[j >> 1] = pack_bf16x2(g2.x, g2.y);
                }
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tma_aux, stage, n, m_box);
                tma_store_commit();
                tma_store_wait_read();
              }
              __syncwarp();
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                *reinterpret_cast<uint4*>(srow + ((q ^ (lane & 7)) << 4)) =
                    make_uint4(act0[q * 4], act0[q * 4 + 1], act0[q * 4 + 2], act0[q * 4 + 3]);
                *reinterpret_cast<uint4*>(srow + (((4 + q) ^ (lane & 7)) << 4)) =
                    make_uint4(act1[q * 4], act1[q * 4 + 1], act1[q * 4 + 2], act1[q * 4 + 3]);
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_3d(&tma_d, stage, n, m_box, batch);
                tma_store_commit();
              }
              continue;
            }
            // operand of this 128-byte-wide chunk through the staging tile (full chunks only; tails read directly)
            const bool grad_epi = (p.epilogue == EPI_DGELU || p.epilogue == EPI_DRELU);
            const bool op_aux = grad_epi && !p.out_fp32 && n + 64 <= p.N;
            const bool op_res = !grad_epi && p.residual != nullptr && p.out_fp32 && n + 32 <= p.N;
            if (op_aux)
              epi_operand_to_stage(stage, reinterpret_cast<const uint8_t*>(p.aux_in + (long long)m_box * p.ld_aux_in + n),
                                   (long long)p.ld_aux_in * 2, p.M - m_box, lane);
            else if (op_res)
              epi_operand_to_stage(stage, reinterpret_cast<const uint8_t*>(p.residual + boff_r + (long long)m_box * p.ldr + n),
                                   (long long)p.ldr * 4, p.M - m_box, lane);
#pragma unroll 1
            for (int hh = 0; hh < nh; ++hh) {
              float f[32];
              const int nn = n + 32 * hh;
              epi_load(p, trow + col0 + 32 * hh, nn, split, f);
              if (op_aux) epi_act_grad_smem(p, srow, lane, hh, f);
              else epi_act(p, row, row_ok, nn, f);
              if (op_res) epi_residual_smem(srow, lane, f);
              else epi_residual(p, boff_r, row, row_ok, nn, f);
              if (hh == 0 && !op_aux && !op_res) {  // the previous store has finished reading the staging tile
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
              }
              if (p.out_fp32) stage_f32(srow, lane, f);
              else stage_bf16(srow, lane, hh, f);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&tma_d, stage, n, m_box, batch);
              tma_store_commit();
            }
          }
        } else if (kb1 > kb0) {
          // ---- split-K: reduce the fp32 partial tile into D (zeroed or accumulating) with vector red.add
#pragma unroll 1
          for (int c = 0; c < kColsPerHalf; c += 32) {
            const int col0 = half * kColsPerHalf + c;
            const int n = n0 + col0;
            if (n >= p.N) break;
            float f[32];
            epi_load(p, trow + col0, n, split, f);
            if (row_ok) {
              float* d = reinterpret_cast<float*>(p.D) + boff_d + (long long)row * p.ldd + n;
              if (n + 32 <= p.N) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) red_add_f32x4(d + j, f[j], f[j + 1], f[j + 2], f[j + 3]);
              } else {
                for (int j = 0; j < 32 && n + j < p.N; ++j) atomicAdd(d + j, f[j]);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
        continue;
      }
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
        if (p.bias != nullptr && split == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full_chunk || n + j < p.N) f[j] += __ldg(p.bias + n + j);
        }
        if (splits > 1) {
          // split-K: reduce the partial tile into fp32 D (zeroed or accumulating); no other epilogue applies
          if (row_ok && kb1 > kb0) {
            float* d = reinterpret_cast<float*>(p.D) + boff_d + (long long)row * p.ldd + n;
            if (full_chunk) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) red_add_f32x4(d + j, f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
              for (int j = 0; j < 32 && n + j < p.N; ++j) atomicAdd(d + j, f[j]);
            }
          }
          continue;
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
    if (BN >= 128 && lane == 0) tma_store_wait_all();  // bulk stores of this thread are complete before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no CTA exits while a peer may still multicast into its smem / barriers
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

int make_tmap_bf16_chunks(CUtensorMap* map, const void* base, uint64_t mn, uint64_t rows, uint64_t batch,
                          uint64_t pitch_elems, uint64_t batch_pitch_elems, uint32_t box_chunks) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return S3D_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch_elems * 2) % 16 != 0 || (batch_pitch_elems * 2) % 16 != 0 ||
      mn % 64 != 0)
    return S3D_ERR_ALIGNMENT;
  cuuint64_t dims[4] = {64, rows, mn / 64, batch};
  cuuint64_t strides[3] = {pitch_elems * 2, 128, batch_pitch_elems * 2};
  cuuint32_t box[4] = {64, 64, box_chunks, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? S3D_OK : S3D_ERR_DRIVER;
}

int make_tmap_bf16_panel(CUtensorMap* map, const void* base, uint64_t rows, uint64_t panels, uint64_t batch,
                         uint32_t box_rows, uint32_t box_panels) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return S3D_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 1023) != 0) return S3D_ERR_ALIGNMENT;
  cuuint64_t dims[4] = {64, rows, panels, batch};
  cuuint64_t strides[3] = {128, rows * 128, panels * rows * 128};
  cuuint32_t box[4] = {64, box_rows, box_panels, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? S3D_OK : S3D_ERR_DRIVER;
}

// rank-2 / rank-3 tensor map over the output (bf16 or fp32), box {box_inner, 32 rows}, 128B swizzle (staged epilogue)
static int make_tmap_out(CUtensorMap* map, const void* base, int is_f32, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                         int batch, uint64_t batch_pitch_elems, bool rank3) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return S3D_ERR_DRIVER;
  const uint64_t es = is_f32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch_elems * es) % 16 != 0) return S3D_ERR_ALIGNMENT;
  const cuuint32_t box_inner = is_f32 ? 32 : 64;  // 128 bytes
  cuuint64_t dims[3] = {inner, outer, (cuuint64_t)(batch > 1 ? batch : 1)};
  cuuint64_t strides[2] = {pitch_elems * es, batch_pitch_elems * es};
  cuuint32_t box[3] = {box_inner, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const int rank = rank3 ? 3 : 2;
  if (rank == 3 && (batch_pitch_elems * es) % 16 != 0) return S3D_ERR_ALIGNMENT;
  CUresult r = enc(map, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? S3D_OK : S3D_ERR_DRIVER;
}

bool pdl_enabled() {
  static const bool on = []() { const char* v = getenv("S3D_PDL"); return v == nullptr || v[0] != '0'; }();
  return on;
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

template <int BN, int A_MN, int B_MN, int CL, int EPW>
static int launch_gemm(const GemmArgs& g, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, EPW>;
  CUtensorMap ta, tb;
  int rc;
  const GemmParams& p = g.p;
  constexpr int kBBoxRows = BN / CL;
  // extent of an operand's batch dimension under the coordinate mapping of GemmParams
  auto bdim = [&](int mul) { return (uint64_t)((g.batch / p.batch_inner - 1) * mul + p.batch_inner); };
  if ((p.a_panel == 1 && !A_MN) || (p.a_panel == 2 && A_MN)) return S3D_ERR_UNSUPPORTED;
  {
    // rank-3 maps always (see the producer): without batches the third extent is 1 and its stride is the matrix size
    const uint64_t ba = g.batch > 1 ? bdim(p.bmul_a) : 1, bb = g.batch > 1 ? bdim(p.bmul_b) : 1;
    const uint64_t sa = g.batch > 1 ? (uint64_t)g.batch_stride_a : (uint64_t)g.lda * (uint64_t)(A_MN ? p.K : p.M);
    const uint64_t sb = g.batch > 1 ? (uint64_t)g.batch_stride_b : (uint64_t)g.ldb * (uint64_t)(B_MN ? p.K : p.N);
    if (A_MN && p.a_panel) rc = make_tmap_bf16_panel(&ta, g.A, p.K, (uint64_t)(p.M + 63) / 64, ba, 64, BM / 64);
    else if (p.a_panel == 2) rc = make_tmap_bf16_panel(&ta, g.A, p.M, (uint64_t)(p.K + 63) / 64, ba, BM, 1);
    else if (A_MN && p.chunk_a) rc = make_tmap_bf16_chunks(&ta, g.A, p.M, p.K, ba, g.lda, sa, BM / 64);
    else if (A_MN) rc = make_tmap_bf16_3d(&ta, g.A, p.M, p.K, ba, g.lda, sa, 64, 64);
    else rc = make_tmap_bf16_3d(&ta, g.A, p.K, p.M, ba, g.lda, sa, 64, BM);
    if (rc) return rc;
    if (B_MN && p.chunk_b) rc = make_tmap_bf16_chunks(&tb, g.B, p.N, p.K, bb, g.ldb, sb, BN / 64 / CL);
    else if (B_MN) rc = make_tmap_bf16_3d(&tb, g.B, p.N, p.K, bb, g.ldb, sb, 64, 64);
    else rc = make_tmap_bf16_3d(&tb, g.B, p.K, p.N, bb, g.ldb, sb, 64, kBBoxRows);
    if (rc) return rc;
  }
  CUtensorMap td, taux;
  memset(&td, 0, sizeof(td));
  memset(&taux, 0, sizeof(taux));
  if (BN >= 128 && p.splits == 1) {
    rc = make_tmap_out(&td, p.D, p.out_fp32, (uint64_t)p.N, (uint64_t)p.M, (uint64_t)p.ldd,
                       g.batch > 1 ? (int)bdim(p.bmul_d) : 1,
                       g.batch > 1 ? (uint64_t)p.batch_stride_d : (uint64_t)p.ldd * (uint64_t)p.M, true);
    if (rc) return rc;
    if (p.aux_out != nullptr) {
      if (g.batch > 1) return S3D_ERR_UNSUPPORTED;
      rc = make_tmap_out(&taux, p.aux_out, 0, (uint64_t)p.N, (uint64_t)p.M, (uint64_t)p.ld_aux_out, 1, 0, false);
      if (rc) return rc;
    }
  }
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, CL, EPW>;
  static bool attr_set = false;
  if (!attr_set) {
    S3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int num_m = (p.M + BM - 1) / BM;
  const int nb = g.batch > 1 ? g.batch : 1;
  const long long items = (long long)((num_m + CL - 1) / CL) * ((p.N + BN - 1) / BN) * p.splits * nb;
  if (items > 0x7fffffffLL) return S3D_ERR_BAD_SHAPE;
  const int sms = num_sms();
  int max_clusters = sms / CL;
  if (max_clusters < 1) max_clusters = 1;
  const int clusters = items < max_clusters ? (int)items : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CL, 1, 1);
  cfg.blockDim = dim3(Cfg::kThreads, 1, 1);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see common.cuh: PDL
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  S3D_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb, td, taux, p));
  return S3D_OK;
}

template <int BN, int CL>
static int dispatch_major(const GemmArgs& g, cudaStream_t s) {
  // 16 epilogue warps for the activation epilogues (forward K-major x K-major, dX K-major x MN-major), see GemmCfg;
  // S3D_GEMM_EPW=2 / 4 forces the 8- / 16-warp epilogue wherever it is instantiated
  static const int forced = []() { const char* v = getenv("S3D_GEMM_EPW"); return v == nullptr ? 0 : atoi(v); }();
  // measured on the cfg3 MLP shapes: dGELU 1002 -> 968 us with 16 warps; the GELU GEMM (two output tiles, 2.3 GB of
  // writes in 0.9 ms: DRAM-write bound) and the plain epilogues are faster with 8 warps and the fourth operand stage
  const bool act = (g.p.epilogue == EPI_DGELU || g.p.epilogue == EPI_DRELU) && g.p.splits == 1;
  // BN = 256 only: a 16-warp part must span at least one 128-byte staging row (64 bf16 columns = BN / 4)
  const bool heavy = BN == 256 && g.a_mn == 0 && (forced == 4 || (forced != 2 && act));
  if (g.a_mn == 0 && g.b_mn == 0) {
    if constexpr (BN == 256) { if (heavy) return launch_gemm<BN, 0, 0, CL, 4>(g, s); }
    return launch_gemm<BN, 0, 0, CL, 2>(g, s);
  }
  if (g.a_mn == 1 && g.b_mn == 1) return launch_gemm<BN, 1, 1, CL, 2>(g, s);
  if (g.a_mn == 0 && g.b_mn == 1) {
    if constexpr (BN == 256) { if (heavy) return launch_gemm<BN, 0, 1, CL, 4>(g, s); }
    return launch_gemm<BN, 0, 1, CL, 2>(g, s);
  }
  return launch_gemm<BN, 1, 0, CL, 2>(g, s);
}

static unsigned env_u32(const char* name, unsigned dflt) {
  const char* v = getenv(name);
  return v ? (unsigned)strtoul(v, nullptr, 0) : dflt;
}

int gemm_bf16(const GemmArgs& g_in, cudaStream_t stream) {
  static const unsigned mn_lbo = env_u32("S3D_DBG_MN_LBO", 8192), mn_sbo = env_u32("S3D_DBG_MN_SBO", 1024);
  static const unsigned k_lbo = env_u32("S3D_DBG_K_LBO", 16), k_sbo = env_u32("S3D_DBG_K_SBO", 1024);
  static const int env_cluster = (int)env_u32("S3D_GEMM_CLUSTER", 0);
  GemmArgs g = g_in;
  g.p.mn_lbo = mn_lbo;
  g.p.mn_sbo = mn_sbo;
  g.p.k_lbo = k_lbo;
  g.p.k_sbo = k_sbo;
  GemmParams& p = g.p;
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return S3D_ERR_BAD_SHAPE;
  p.batch = g.batch > 1 ? g.batch : 1;
  p.batch_inner = g.batch_inner > 0 ? g.batch_inner : 1;
  if (p.batch % p.batch_inner != 0) return S3D_ERR_BAD_SHAPE;
  p.bmul_a = g.batch_inner > 0 ? g.bmul_a : 1;
  p.bmul_b = g.batch_inner > 0 ? g.bmul_b : 1;
  p.bmul_d = g.batch_inner > 0 ? g.bmul_d : 1;
  if (p.a_panel) p.batched = 1;  // panel operands are addressed through rank-4 / rank-3 maps even for a single batch
  static const bool no_chunks = getenv("S3D_GEMM_NO_CHUNK_VIEW") != nullptr;
  p.chunk_a = (g.a_mn && (p.a_panel == 1 || (!no_chunks && p.M % 64 == 0))) ? 1 : 0;
  p.chunk_b = (g.b_mn && !no_chunks && p.N % 64 == 0) ? 1 : 0;
  if (p.batch > 1 && !p.batched) return S3D_ERR_BAD_SHAPE;
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
  const int sms = num_sms();
  const int nb = g.batch > 1 ? g.batch : 1;
  const int num_m = (p.M + BM - 1) / BM;
  const int num_kb = (p.K + BK - 1) / BK;
  int bn = g.force_bn;
  if (bn == 0) {
    // BN=256 halves the smem/L2 operand traffic per flop; prefer it whenever the problem still offers enough tiles
    // (split-K can recover parallelism for fp32 outputs).
    const long long t256 = (long long)num_m * ((p.N + 255) / 256) * nb;
    const long long t128 = (long long)num_m * ((p.N + 127) / 128) * nb;
    if (p.N > 128 && (t256 >= sms || (p.N % 256 == 0 && t256 * 4 >= sms))) bn = 256;
    else if (p.N > 64 && (t128 >= sms / 2 || p.N % 128 == 0)) bn = 128;
    else bn = (p.N > 64 && t128 * 2 > sms) ? 128 : 64;
  }
  // Small problems (at most one tile per SM even with 256-wide tiles: the cfg2 / point-model shapes) are latency bound,
  // not throughput bound (in-graph probe, profiles/r01_small_gemm_variants_in_graph.log): a launch costs 6-11 us whatever
  // the variant, and the best tile width is the one that spreads the work over the most SMs while staying in a single
  // round (e.g. M=1664: N=1152 -> BN128 = 117 tiles, N=1536 -> BN256 = 78 tiles, N=384 -> BN64 = 78 tiles). Long-K
  // weight gradients are the exception: split-K supplies the parallelism and a narrow tile only re-reads the A operand.
  bool small_latency_bound = false;
  static const bool small_rule = []() { const char* v = getenv("S3D_GEMM_SMALL"); return v == nullptr || v[0] != '0'; }();
  if (small_rule && g.force_bn == 0 && (long long)num_m * ((p.N + 255) / 256) * nb <= sms) {
    auto tiles_of = [&](int b) { return (long long)num_m * ((p.N + b - 1) / b) * nb; };
    auto waste_ok = [&](int b) { const int nn = (p.N + b - 1) / b * b; return (nn - p.N) * 4 <= nn; };
    if (g.a_mn && g.b_mn && num_kb >= 64) {
      bn = waste_ok(256) ? 256 : (waste_ok(128) ? 128 : 64);
    } else {
      long long best = -1;
      for (int b : {256, 128, 64}) {
        if (b != 64 && p.N <= b / 2) continue;
        const long long t = tiles_of(b);
        if (t <= sms && t > best) { best = t; bn = b; }
      }
      small_latency_bound = true;
    }
  }
  if (bn != 64 && bn != 128 && bn != 256) return S3D_ERR_UNSUPPORTED;
  const long long tiles = (long long)num_m * ((p.N + bn - 1) / bn) * nb;
  // cluster size along M (B-tile multicast): needs >= 2 m-tiles and whole 64-wide chunks of B per CTA
  int cl = g.force_cluster > 0 ? g.force_cluster : (env_cluster > 0 ? env_cluster : 2);
  if (cl > bn / 64) cl = bn / 64;
  if (num_m < 2 || cl < 1) cl = 1;
  if (cl == 4 && (bn != 256 || num_m < 4)) cl = 2;
  if (cl == 3 || cl > 4) cl = 2;
  if (small_latency_bound && g.force_cluster == 0 && env_cluster == 0) cl = 1;  // single round: a cluster launch only adds latency
  // split-K: plain fp32 output (optionally accumulating onto itself), too few tiles to fill the machine, long K
  const bool split_ok = p.out_fp32 && p.epilogue == EPI_NONE && p.aux_out == nullptr &&
                        (p.residual == nullptr || p.residual == p.D) && nb == 1;
  int splits = 1;
  if (g.force_splits > 0) splits = split_ok ? g.force_splits : 1;
  else if (split_ok && g.a_mn && g.b_mn && tiles * 2 <= sms && num_kb >= 8) {
    // automatic split-K only for weight-gradient GEMMs (both operands MN-major, K = tokens): red.add makes the fp32
    // summation order non-deterministic, which forward / activation-gradient GEMMs must not be.
    // The persistent kernel walks work items in rounds of `slots` clusters: pick the smallest split factor whose last
    // round is (nearly) full -- e.g. 27 cluster tiles on 74 slots: 5 splits = 135 items = 2 rounds at 91 %, 6 splits =
    // 3 rounds at 73 %, 8 splits = 216 items = 3 rounds at 97 %.
    const long long ctiles = (long long)((num_m + cl - 1) / cl) * ((p.N + bn - 1) / bn);
    const long long slots = sms / cl > 0 ? sms / cl : 1;
    int smax = (int)((4 * slots + ctiles - 1) / ctiles);
    if (smax > num_kb / 4) smax = num_kb / 4;
    if (smax < 1) smax = 1;
    double best = 0.0;
    for (int sp = 1; sp <= smax; ++sp) {
      const long long items = ctiles * sp;
      const long long rounds = (items + slots - 1) / slots;
      const double eff = (double)items / (double)(rounds * slots);
      if (eff > best) best = eff;
    }
    for (int sp = 1; sp <= smax; ++sp) {
      const long long items = ctiles * sp;
      const long long rounds = (items + slots - 1) / slots;
      const double eff = (double)items / (double)(rounds * slots);
      if (eff >= 0.97 * best) { splits = sp; break; }
    }
  }
  if (splits > num_kb) splits = num_kb;
  p.splits = splits;
  if (splits > 1) {
    if (p.residual == nullptr) {
      // partial sums are reduced with red.add: start from zero (row by row when D is a strided view)
      if (p.ldd == p.N) S3D_CUDA_OK(cudaMemsetAsync(p.D, 0, sizeof(float) * (size_t)p.M * p.N, stream));
      else S3D_CUDA_OK(cudaMemset2DAsync(p.D, sizeof(float) * p.ldd, 0, sizeof(float) * p.N, p.M, stream));
    }
    p.residual = nullptr;  // accumulate-in-place is what the reduction does anyway
  }
  switch (bn) {
    case 256:
      if (cl == 4) return dispatch_major<256, 4>(g, stream);
      if (cl == 2) return dispatch_major<256, 2>(g, stream);
      return dispatch_major<256, 1>(g, stream);
    case 128:
      if (cl == 2) return dispatch_major<128, 2>(g, stream);
      return dispatch_major<128, 1>(g, stream);
    default: return dispatch_major<64, 1>(g, stream);
  }
}

}  // namespace s3d
