// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers.
// Everything here is inline PTX; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace s3d {

// ----------------------------------------------------------------------------------------------
// status codes shared with include/s3d_b200.h (negative = argument error, positive = cudaError_t)
// ----------------------------------------------------------------------------------------------
enum : int {
  S3D_OK = 0,
  S3D_ERR_BAD_SHAPE = -1,
  S3D_ERR_UNSUPPORTED = -2,
  S3D_ERR_ALIGNMENT = -3,
  S3D_ERR_NULL = -4,
  S3D_ERR_DRIVER = -5,
  S3D_ERR_WORKSPACE = -6,
};

#define S3D_CUDA_OK(expr)                   \
  do {                                      \
    cudaError_t _e = (expr);                \
    if (_e != cudaSuccess) return (int)_e;  \
  } while (0)

#define S3D_LAUNCH_OK()                     \
  do {                                      \
    cudaError_t _e = cudaGetLastError();    \
    if (_e != cudaSuccess) return (int)_e;  \
  } while (0)

// ----------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL). The small configurations are launch-latency bound (cfg2: ~280 kernels of
// 5-10 us per step), so consecutive kernels of ours are launched with programmaticStreamSerialization: the next grid
// is scheduled, and runs its prologue (barrier init, TMEM allocation, descriptor prefetch), while the previous grid
// drains. Contract: every kernel launched through launch_pdl() calls pdl_prologue() (or pdl_trigger() + pdl_wait())
// BEFORE its first global-memory access; griddepcontrol.wait returns once all prerequisite grids have completed and
// their writes are visible, so data dependencies (RAW / WAR / WAW, transitively) are exactly those of a plain stream.
// S3D_PDL=0 in the environment launches everything without the attribute (the instructions are then no-ops).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_trigger();
  pdl_wait();
}
bool pdl_enabled();  // gemm_tcgen05.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------- mbarrier -------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the hardware parks the warp until the phase completes (or the hint expires)
// instead of returning after a few dozen cycles, so a waiting producer / MMA warp stops competing for the issue slots
// of the element-wise warps on its scheduler (measured on the flash forward kernel: ~20 % of all issued instructions
// were spin-loop bookkeeping).
__device__ __forceinline__ bool mbar_try_wait_parked(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (context error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait_parked(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}

// ---------------------------------------------- TMA ----------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               int c2, int c3, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
      "%4, %5, %6}], [%2], %7;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
      : "memory");
}

// multicast variants: the box lands at the same smem offset in every CTA of `mask`, each CTA's mbarrier gets the bytes
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
      "%4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               int c2, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
      "%4, %5}], [%2], %6;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}

// ------------------------------------------- clusters -------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// -------------------------------------------- tcgen05 --------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tcgen05.commit: the mbarrier receives one arrive when all previously issued MMAs of this thread retire.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// multicast commit: one arrive on the barrier at the same smem offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// fp32 vector reduction to global memory (split-K partial sums)
__device__ __forceinline__ void red_add_f32x4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16/fp16 inputs, fp32 accumulate.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from TMEM (used by the attention kernels: P stays in tensor memory).
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 32 lanes x 32 columns of 32-bit: thread i of the warp receives row (lane base + i), 32 consecutive columns.
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

// ---------------------------------------- UMMA descriptors ----------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B (layout_type 2), descriptor version 1 (sm_100).
//   K-major tile  [rows][64 bf16]: 128-byte rows, 8-row swizzle atoms 1024 B apart  -> SBO = 1024, LBO unused.
//   MN-major tile [k][64 bf16]   : 128-byte k-rows, 8-k atoms 1024 B apart (SBO), 64-element MN chunks LBO apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

// Split form of the same descriptor: the high word is a compile-time constant and the low word is
// (start address >> 4) | (LBO >> 4) << 16, so stepping along K inside the MMA issue loop is one 32-bit add.
__host__ __device__ __forceinline__ constexpr uint32_t smem_desc_hi_sw128(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
}
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ void umma_f16_ss2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                          // D format = F32
         | (1u << 7)                        // A format = BF16
         | (1u << 10)                       // B format = BF16
         | ((uint32_t)a_mn_major << 15)     // A major (0 = K)
         | ((uint32_t)b_mn_major << 16)     // B major (0 = K)
         | ((uint32_t)(N >> 3) << 17)       // N / 8
         | ((uint32_t)(M >> 4) << 24);      // M / 16
}

// --------------------------------------------- misc ---------------------------------------------
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax polynomial, rel. err ~1e-4 -- below bf16 resolution):
// used for half of the softmax exponentials so the MUFU pipe (16 ex2/clk/SM) is not the attention bottleneck.
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.0f);
  const float fl = floorf(x);
  const float fr = x - fl;
  float p = fmaf(0.0771190f, fr, 0.2275643f);
  p = fmaf(p, fr, 0.6951461f);
  p = fmaf(p, fr, 1.0f);
  return __int_as_float(__float_as_int(p) + (static_cast<int>(fl) << 23));
}

// ------------------------------------------ packed fp32 pairs ------------------------------------------
// Blackwell executes fma / add / mul on two fp32 lanes per instruction (FFMA2 / FADD2 / FMUL2): the element-wise warps
// of the attention kernels are instruction-issue bound, so their scale / subtract / accumulate steps run packed.
#ifdef __CUDACC__
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n"
      ".reg .b64 ra, rb, rc, rd;\n"
      "mov.b64 ra, {%2, %3};\n"
      "mov.b64 rb, {%4, %5};\n"
      "mov.b64 rc, {%6, %7};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\n"
      "mov.b64 {%0, %1}, rd;\n"
      "}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n"
      ".reg .b64 ra, rb, rd;\n"
      "mov.b64 ra, {%2, %3};\n"
      "mov.b64 rb, {%4, %5};\n"
      "add.rn.f32x2 rd, ra, rb;\n"
      "mov.b64 {%0, %1}, rd;\n"
      "}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{\n"
      ".reg .b64 ra, rb, rd;\n"
      "mov.b64 ra, {%2, %3};\n"
      "mov.b64 rb, {%4, %5};\n"
      "mul.rn.f32x2 rd, ra, rb;\n"
      "mov.b64 {%0, %1}, rd;\n"
      "}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
#endif

// Branch-free erf (Abramowitz & Stegun 7.1.26, |abs err| <= 1.5e-7): keeps the fused GELU epilogues short -- the
// library erff() expands to ~40 instructions with two branches per element, which made every GEMM kernel > 150 KB of SASS.
// GELU(x) = x * Phi(x) with the exact (erf) Gaussian CDF, evaluated as an odd Chebyshev-fitted polynomial:
//   Phi(x) - 1/2  ~  u * P(u^2),  u = clamp(x, -4, 4) / 4,   degree 17, |error| <= 5e-6 in fp32 (far below bf16 resolution);
//   GELU'(x) - 1/2 = Phi(x) + x phi(x) - 1/2  ~  u * Q(u^2),  u = clamp(x, -4.5, 4.5) / 4.5,  degree 21, |error| <= 1.3e-4.
// ~14 FMA-pipe instructions per element and no MUFU, instead of ~35 with erf + exp: the fused GEMM epilogues are
// instruction-issue bound (ncu: 63 % issue-slot utilisation, 40 instructions per output element before this change).
__device__ __forceinline__ float gelu_cdf(float x) {
  const float u = fminf(fmaxf(x, -4.0f), 4.0f) * 0.25f;
  const float u2 = u * u;
  float p = 1.340839184e+00f;
  p = fmaf(p, u2, -7.331169602e+00f);
  p = fmaf(p, u2, 1.789781136e+01f);
  p = fmaf(p, u2, -2.609703610e+01f);
  p = fmaf(p, u2, 2.576648281e+01f);
  p = fmaf(p, u2, -1.852975207e+01f);
  p = fmaf(p, u2, 1.010684673e+01f);
  p = fmaf(p, u2, -4.249730180e+00f);
  p = fmaf(p, u2, 1.595679461e+00f);
  return fmaf(p, u, 0.5f);
}
__device__ __forceinline__ float gelu_erf(float x) { return x * gelu_cdf(x); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float u = fminf(fmaxf(x, -4.5f), 4.5f) * (1.0f / 4.5f);
  const float u2 = u * u;
  float p = 5.250829664e+01f;
  p = fmaf(p, u2, -3.295807064e+02f);
  p = fmaf(p, u2, 9.265840972e+02f);
  p = fmaf(p, u2, -1.548898734e+03f);
  p = fmaf(p, u2, 1.726300807e+03f);
  p = fmaf(p, u2, -1.364789792e+03f);
  p = fmaf(p, u2, 7.934743716e+02f);
  p = fmaf(p, u2, -3.440685097e+02f);
  p = fmaf(p, u2, 1.095854887e+02f);
  p = fmaf(p, u2, -2.420539351e+01f);
  p = fmaf(p, u2, 3.590152229e+00f);
  return fmaf(p, u, 0.5f);
}
#ifdef __CUDACC__
// The same two polynomials on PACKED fp32 pairs (FFMA2 / FMUL2), with the 1/4 and 1/4.5 argument scalings folded into
// the coefficients: 7.5 / 8.5 issued instructions per element instead of 14 / 16. The GELU / dGELU GEMM epilogues are
// instruction-issue bound (ncu: tensor pipe 48-52 % on those two shapes against 69-88 % on the plain ones).
__device__ __forceinline__ float2 bcast2(float c) { return make_float2(c, c); }
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 t = make_float2(fminf(fmaxf(x.x, -4.0f), 4.0f), fminf(fmaxf(x.y, -4.0f), 4.0f));
  const float2 t2 = fmul2(t, t);
  // c_k / 4^(2k+1) of gelu_cdf's coefficients
  float2 p = bcast2(1.340839184e+00f / 17179869184.0f);
  p = ffma2(p, t2, bcast2(-7.331169602e+00f / 1073741824.0f));
  p = ffma2(p, t2, bcast2(1.789781136e+01f / 67108864.0f));
  p = ffma2(p, t2, bcast2(-2.609703610e+01f / 4194304.0f));
  p = ffma2(p, t2, bcast2(2.576648281e+01f / 262144.0f));
  p = ffma2(p, t2, bcast2(-1.852975207e+01f / 16384.0f));
  p = ffma2(p, t2, bcast2(1.010684673e+01f / 1024.0f));
  p = ffma2(p, t2, bcast2(-4.249730180e+00f / 64.0f));
  p = ffma2(p, t2, bcast2(1.595679461e+00f / 4.0f));
  return fmul2(x, ffma2(p, t, bcast2(0.5f)));
}
// t already clamped to [-4.5, 4.5] (the GEMM epilogue clamps the packed bf16 pre-activations with two bf16x2 min / max
// instructions per pair instead of four fp32 ones)
__device__ __forceinline__ float2 gelu_erf_grad2_clamped(float2 t);
__device__ __forceinline__ float2 gelu_erf_grad2(float2 x) {
  return gelu_erf_grad2_clamped(make_float2(fminf(fmaxf(x.x, -4.5f), 4.5f), fminf(fmaxf(x.y, -4.5f), 4.5f)));
}
__device__ __forceinline__ uint32_t clamp_bf16x2_4p5(uint32_t v) {
  uint32_t d;
  // 0x4090 = bf16(4.5), 0xC090 = bf16(-4.5)
  asm("{\n"
      ".reg .b32 t;\n"
      "min.bf16x2 t, %1, %2;\n"
      "max.bf16x2 %0, t, %3;\n"
      "}\n"
      : "=r"(d)
      : "r"(v), "r"(0x40904090u), "r"(0xC090C090u));
  return d;
}
__device__ __forceinline__ float2 gelu_erf_grad2_clamped(float2 t) {
  const float2 u = fmul2(t, bcast2(1.0f / 4.5f));  // coefficients of u^21 would underflow the fp32 range if folded
  const float2 u2 = fmul2(u, u);
  float2 p = bcast2(5.250829664e+01f);
  p = ffma2(p, u2, bcast2(-3.295807064e+02f));
  p = ffma2(p, u2, bcast2(9.265840972e+02f));
  p = ffma2(p, u2, bcast2(-1.548898734e+03f));
  p = ffma2(p, u2, bcast2(1.726300807e+03f));
  p = ffma2(p, u2, bcast2(-1.364789792e+03f));
  p = ffma2(p, u2, bcast2(7.934743716e+02f));
  p = ffma2(p, u2, bcast2(-3.440685097e+02f));
  p = ffma2(p, u2, bcast2(1.095854887e+02f));
  p = ffma2(p, u2, bcast2(-2.420539351e+01f));
  p = ffma2(p, u2, bcast2(3.590152229e+00f));
  return ffma2(p, u, bcast2(0.5f));
}
#endif

// Counter-based dropout mask (nn.Dropout / attention-probability dropout of the group_embed layer, reference
// vit_3d_2d_pretrain.py:381: nn.TransformerEncoderLayer's default p = 0.1, active in train()). Stateless: backward
// regenerates the forward mask from (seed, site, row, col).
//   site seed  ss      = mix(seed ^ site * C_site ^ const)
//   row key    rk(row) = mix(ss ^ row * C_row)                                  (once per row)
//   pair word  z(row, cp) = fold(fold(rk + cp * C_col, M1), M2) & 0x3FFF3FFF,   fold(x, m) = lo32(x * m) ^ hi32(x * m)
//   element (row, col) draws the 14-bit value r = col odd ? z >> 16 : z & 0x3FFF of pair cp = col >> 1 and is KEPT when
//   r >= thresh14 = round(p * 16384); kept values are scaled by 1 / (1 - thresh14 / 16384).
// The pair word costs 5 integer instructions (IADD, 2 x IMAD.WIDE, 2 x LOP3). Both 14-bit halves are bit patterns of
// finite non-negative fp16 numbers (< 2.0), whose order equals their integer order, so ONE half2 comparison
// (set.ge.u32.f16x2 -> HSET2) turns z into the 0xFFFF / 0x0000 AND-masks of a packed bf16x2 pair, and setp.ge.f16x2
// yields both predicates at once. (One fold round leaves lattice structure -- P(drop, drop) at column distance 2 is 2x
// off -- two rounds pass the pair / row-sum / column-sum statistics checked in tests/test_dropout_hash.py.)
constexpr uint32_t kDropRowMul = 0x9E3779B1u, kDropColMul = 0x85EBCA77u, kDropSiteMul = 0x632BE5ABu;
constexpr uint32_t kDropM1 = 0xD6E8FEB9u, kDropM2 = 0xCA6B1B35u;
__host__ __device__ __forceinline__ uint32_t drop_mix(uint32_t x) {
  x *= 0x2C1B3C6Du;
  x ^= x >> 15;
  x *= 0x297A2D39u;
  x ^= x >> 15;
  return x;
}
__host__ __device__ __forceinline__ uint32_t drop_site_seed(uint32_t seed, uint32_t site) {
  return drop_mix(seed ^ (site * kDropSiteMul) ^ 0xA511E9B3u);
}
__host__ __device__ __forceinline__ uint32_t drop_rowkey(uint32_t site_seed, uint32_t row) {
  return drop_mix(site_seed ^ (row * kDropRowMul));
}
__host__ __device__ __forceinline__ uint32_t drop_fold(uint32_t x, uint32_t m) {
  const unsigned long long w = (unsigned long long)x * m;
  return (uint32_t)w ^ (uint32_t)(w >> 32);
}
// y = rowkey + cp * kDropColMul (callers step y by kDropColMul from pair to pair)
__host__ __device__ __forceinline__ uint32_t drop_word(uint32_t y) {
  return drop_fold(drop_fold(y, kDropM1), kDropM2) & 0x3FFF3FFFu;
}
__host__ __device__ __forceinline__ uint32_t drop_pair(uint32_t site_seed, uint32_t row, uint32_t col_pair) {
  return drop_word(drop_rowkey(site_seed, row) + col_pair * kDropColMul);
}
__host__ __device__ __forceinline__ bool drop_keep(uint32_t site_seed, uint32_t row, uint32_t col, uint32_t thresh14) {
  const uint32_t z = drop_pair(site_seed, row, col >> 1);
  return ((col & 1u) ? (z >> 16) : (z & 0xffffu)) >= thresh14;
}
__host__ __device__ __forceinline__ uint32_t drop_thresh14(float p) {
  const float t = p * 16384.0f + 0.5f;
  return t <= 0.f ? 0u : (t >= 16383.f ? 16383u : (uint32_t)t);
}
__host__ __device__ __forceinline__ float drop_keep_scale(uint32_t thresh14) {
  return 1.0f / (1.0f - (float)thresh14 / 16384.0f);
}
#ifdef __CUDACC__
// 0xFFFF / 0x0000 per 16-bit half of z: keep-masks of the pair (thresh2 = thresh14 in both halves)
__device__ __forceinline__ uint32_t drop_andmask(uint32_t z, uint32_t thresh2) {
  uint32_t d;
  asm("set.ge.u32.f16x2 %0, %1, %2;" : "=r"(d) : "r"(z), "r"(thresh2));
  return d;
}
// zero a (even column) / b (odd column) when dropped
__device__ __forceinline__ void drop_zero2(float& a, float& b, uint32_t z, uint32_t thresh2) {
  asm("{\n"
      ".reg .pred p, q;\n"
      "setp.ge.f16x2 p|q, %2, %3;\n"
      "@!p mov.f32 %0, 0f00000000;\n"
      "@!q mov.f32 %1, 0f00000000;\n"
      "}\n"
      : "+f"(a), "+f"(b)
      : "r"(z), "r"(thresh2));
}
#endif

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Host: resolves cuTensorMapEncodeTiled through the runtime (no link-time libcuda dependency).
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();

// rank-2 bf16 tensor map: dims {inner, outer}, row pitch in elements, box {box_inner, box_outer}, 128B swizzle.
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                      uint32_t box_inner, uint32_t box_outer);
// rank-3 (batched) variant: dims {inner, outer, batch}.
int make_tmap_bf16_3d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t batch,
                      uint64_t pitch_elems, uint64_t batch_pitch_elems, uint32_t box_inner, uint32_t box_outer);

// 64-column panel layout [batch][panel = col / 64][row][64] of a bf16 matrix (every 64 x rows panel contiguous, 128-byte
// rows): dims {64, rows, panels, batch}, box {64, box_rows, 1, 1}, 128B swizzle. Used for the attention backward's
// workspace matrices: a 128 x 64 tile store and a 64 x 64 operand load are single contiguous runs in HBM.
int make_tmap_bf16_panel(CUtensorMap* map, const void* base, uint64_t rows, uint64_t panels, uint64_t batch,
                         uint32_t box_rows, uint32_t box_panels);
// "chunk view" of a row-major bf16 [rows, mn] matrix (mn a multiple of 64) read as an MN-major UMMA operand: dims
// {64, rows, mn / 64, batch} with strides {pitch, 128 B, batch pitch}, box {64, 64, box_chunks, 1}: ONE TMA load brings
// box_chunks 64-wide column chunks of 64 rows, chunk after chunk in shared memory (what the MN-major descriptors expect).
int make_tmap_bf16_chunks(CUtensorMap* map, const void* base, uint64_t mn, uint64_t rows, uint64_t batch,
                          uint64_t pitch_elems, uint64_t batch_pitch_elems, uint32_t box_chunks);

int num_sms();

}  // namespace s3d
