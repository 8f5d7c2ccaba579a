// binvox payload -> dense occupancy grid on the device (the caller-side wire format).
// Reference: utils/binvox_rw.py:117-151 read_as_3d_array: np.repeat(values, counts) over (value, count) byte pairs,
// reshape(dims) in x-z-y order, transpose(0, 2, 1) to x-y-z; data/modelnet40.py:35-45 wraps it as [1, V, V, V].
// The RLE payload (tens of KB per 128^3 model instead of 2 MiB dense, 8 MiB as the reference's int32) is what crosses
// PCIe; the expansion runs at HBM speed:
//   binvox_scan_kernel   : per model, inclusive prefix sum of the run lengths -> run_end[r] (uint32), total voxel count
//   binvox_expand_kernel : one CTA per (model, x-slab): the slab is the contiguous range [x V^2, (x+1) V^2) of the
//                          run-length stream; binary search for its first run, the CTA's threads paint their runs into a
//                          V x V byte plane in shared memory (z-major as stored), then the plane is written transposed
//                          (y-major) with coalesced stores as uint8 / int32 / float.
#include "kernels.h"

namespace s3d {
namespace {

constexpr int kScanThreads = 1024;

// 16 runs (32 payload bytes) per thread and iteration: serial prefix inside the thread, warp + block scan of the thread
// totals, one carry per iteration -> 16384 runs per iteration of the 1024-thread CTA that owns a model.
constexpr int kRunsPerThread = 16;

__global__ void __launch_bounds__(kScanThreads) binvox_scan_kernel(const uint8_t* __restrict__ payload,
                                                                   const long long* __restrict__ offsets,
                                                                   unsigned* __restrict__ run_end,
                                                                   const long long* __restrict__ run_offsets,
                                                                   long long* __restrict__ totals) {
  __shared__ unsigned warp_sums[32];
  __shared__ unsigned carry_s;
  const int b = blockIdx.x;
  const uint8_t* pairs = payload + offsets[b];
  const long long nruns = (offsets[b + 1] - offsets[b]) / 2;
  unsigned* out = run_end + run_offsets[b];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(pairs) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (long long base = 0; base < nruns; base += (long long)kScanThreads * kRunsPerThread) {
    const long long r0 = base + (long long)threadIdx.x * kRunsPerThread;
    unsigned pre[kRunsPerThread];
    unsigned run = 0;
    if (vec_ok && r0 + kRunsPerThread <= nruns) {
      const uint4 u0 = *reinterpret_cast<const uint4*>(pairs + 2 * r0);
      const uint4 u1 = *reinterpret_cast<const uint4*>(pairs + 2 * r0 + 16);
      const unsigned w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {  // each 32-bit word holds two (value, count) pairs: counts are bytes 1 and 3
        run += (w[j] >> 8) & 0xffu;
        pre[2 * j] = run;
        run += (w[j] >> 24) & 0xffu;
        pre[2 * j + 1] = run;
      }
    } else {
#pragma unroll
      for (int j = 0; j < kRunsPerThread; ++j) {
        if (r0 + j < nruns) run += (unsigned)pairs[2 * (r0 + j) + 1];
        pre[j] = run;
      }
    }
    unsigned v = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) warp_sums[warp] = v;
    __syncthreads();
    if (warp == 0) {
      unsigned w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const unsigned excl = (v - run) + (warp > 0 ? warp_sums[warp - 1] : 0u) + carry_s;
    if (vec_ok && r0 + kRunsPerThread <= nruns) {
#pragma unroll
      for (int j = 0; j < kRunsPerThread; j += 4)
        *reinterpret_cast<uint4*>(out + r0 + j) =
            make_uint4(excl + pre[j], excl + pre[j + 1], excl + pre[j + 2], excl + pre[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < kRunsPerThread; ++j)
        if (r0 + j < nruns) out[r0 + j] = excl + pre[j];
    }
    __syncthreads();
    if (threadIdx.x == kScanThreads - 1) carry_s = excl + run;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[b] = (long long)carry_s;
}

template <typename TOut>
__device__ __forceinline__ void store4(TOut* dst, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3);
template <>
__device__ __forceinline__ void store4<uint8_t>(uint8_t* dst, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
  *reinterpret_cast<uint32_t*>(dst) = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
}
template <>
__device__ __forceinline__ void store4<int>(int* dst, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
  *reinterpret_cast<int4*>(dst) = make_int4((int)b0, (int)b1, (int)b2, (int)b3);
}
template <>
__device__ __forceinline__ void store4<float>(float* dst, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
  *reinterpret_cast<float4*>(dst) = make_float4((float)b0, (float)b1, (float)b2, (float)b3);
}

// plane row stride: odd, so that a warp reading one byte per row (the z <-> y transpose) touches 32 different banks
__host__ __device__ __forceinline__ int plane_stride(int V) { return V | 1; }

template <typename TOut>
__global__ void __launch_bounds__(256) binvox_expand_kernel(const uint8_t* __restrict__ payload,
                                                            const long long* __restrict__ offsets,
                                                            const unsigned* __restrict__ run_end,
                                                            const long long* __restrict__ run_offsets,
                                                            TOut* __restrict__ out, int V, int fix_coords) {
  extern __shared__ __align__(16) uint8_t plane[];  // [V][PS] in stream order (row = z, column = y), PS = V | 1
  const int b = blockIdx.y, x = blockIdx.x;
  const int PS = plane_stride(V);
  const uint8_t* pairs = payload + offsets[b];
  const unsigned* ends = run_end + run_offsets[b];
  const int nruns = (int)((offsets[b + 1] - offsets[b]) / 2);
  const unsigned slab = (unsigned)V * (unsigned)V;
  const unsigned lo = (unsigned)x * slab, hi = lo + slab;
  {  // voxels past a short stream read as empty
    const int n16 = (V * PS + 15) / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) reinterpret_cast<uint4*>(plane)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  // first run whose end is > lo: two parallel narrowing rounds (every thread probes one sample of the current range, the
  // CTA keeps the last sample that is still <= lo) and a short serial tail: 3 dependent load latencies instead of ~19
  __shared__ int range_lo;
  int first = 0;
  {
    int l = 0, h = nruns;  // invariant: ends[l - 1] <= lo (or l == 0), answer in [l, h]
    for (int round = 0; round < 2 && h - l > 8; ++round) {
      const int step = (h - l + (int)blockDim.x - 1) / (int)blockDim.x;
      if (threadIdx.x == 0) range_lo = l;
      __syncthreads();
      const int probe = l + (int)threadIdx.x * step;  // ends[probe - 1] <= lo  =>  answer >= probe
      if (probe > l && probe <= h && ends[probe - 1] <= lo) atomicMax(&range_lo, probe);
      __syncthreads();
      const int nl = range_lo;
      h = min(h, nl + step);
      l = nl;
      __syncthreads();
    }
    while (l < h) {
      const int m = (l + h) >> 1;
      if (ends[m] > lo) h = m; else l = m + 1;
    }
    first = l;
  }
  __syncthreads();
  for (int r = first + threadIdx.x; r < nruns; r += blockDim.x) {
    const unsigned e = ends[r];
    const uint2 pr = make_uint2(pairs[2 * r], pairs[2 * r + 1]);
    const unsigned s = e - pr.y;
    if (s >= hi) break;  // runs are ordered: everything this thread would see later is past the slab too
    if (pr.x != 0) {
      const unsigned a = (s > lo ? s : lo) - lo, z = (e < hi ? e : hi) - lo;
      unsigned row = a / (unsigned)V, col = a - row * (unsigned)V;
      for (unsigned i = a; i < z; ++i) {
        plane[row * PS + col] = 1;
        if (++col == (unsigned)V) { col = 0; ++row; }
      }
    }
  }
  __syncthreads();
  // write out: out[b][x][j][k]; fix_coords: (j, k) = (y, z) <- plane[z][y]; otherwise the stream order (z, y) is kept
  TOut* o = out + ((size_t)b * V + x) * slab;
  if ((V & 3) == 0) {
    const int vq = V >> 2;
    for (int g = threadIdx.x; g < (int)(slab >> 2); g += blockDim.x) {
      const int j = g / vq, k = (g - j * vq) << 2;
      uint32_t v0, v1, v2, v3;
      if (fix_coords) {
        const uint8_t* src = plane + k * PS + j;
        v0 = src[0]; v1 = src[PS]; v2 = src[2 * PS]; v3 = src[3 * PS];
      } else {
        const uint8_t* src = plane + j * PS + k;
        v0 = src[0]; v1 = src[1]; v2 = src[2]; v3 = src[3];
      }
      store4<TOut>(o + (size_t)g * 4, v0, v1, v2, v3);
    }
  } else {
    for (int i = threadIdx.x; i < (int)slab; i += blockDim.x) {
      const int j = i / V, k = i - j * V;
      o[i] = (TOut)(fix_coords ? plane[k * PS + j] : plane[j * PS + k]);
    }
  }
}

}  // namespace
}  // namespace s3d

using namespace s3d;

extern "C" {

int s3d_binvox_scan(const uint8_t* payload, const int64_t* offsets, uint32_t* run_end, const int64_t* run_offsets,
                    int64_t* totals, int B, void* stream) {
  if (B <= 0) return S3D_ERR_BAD_SHAPE;
  if (payload == nullptr || offsets == nullptr || run_end == nullptr || run_offsets == nullptr || totals == nullptr)
    return S3D_ERR_NULL;
  binvox_scan_kernel<<<B, kScanThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      payload, reinterpret_cast<const long long*>(offsets), run_end, reinterpret_cast<const long long*>(run_offsets),
      reinterpret_cast<long long*>(totals));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int s3d_binvox_expand(const uint8_t* payload, const int64_t* offsets, const uint32_t* run_end,
                      const int64_t* run_offsets, void* out, int out_dtype, int B, int V, int fix_coords, void* stream) {
  if (B <= 0 || V <= 0) return S3D_ERR_BAD_SHAPE;
  if (V > 384 || B > 65535) return S3D_ERR_UNSUPPORTED;  // V x V byte plane in shared memory
  if (payload == nullptr || offsets == nullptr || run_end == nullptr || run_offsets == nullptr || out == nullptr)
    return S3D_ERR_NULL;
  const size_t smem = (((size_t)V * plane_stride(V) + 15) / 16) * 16;
  const dim3 grid(V, B);
  auto st = reinterpret_cast<cudaStream_t>(stream);
  auto off = reinterpret_cast<const long long*>(offsets);
  auto roff = reinterpret_cast<const long long*>(run_offsets);
#define S3D_BINVOX(T)                                                                                              \
  do {                                                                                                             \
    auto kern = binvox_expand_kernel<T>;                                                                           \
    if (smem > 48 * 1024) S3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, 256, smem, st>>>(payload, off, run_end, roff, reinterpret_cast<T*>(out), V, fix_coords);         \
  } while (0)
  switch (out_dtype) {
    case 0: S3D_BINVOX(float); break;
    case 1: S3D_BINVOX(uint8_t); break;
    case 2: S3D_BINVOX(int); break;
    default: return S3D_ERR_UNSUPPORTED;
  }
#undef S3D_BINVOX
  S3D_LAUNCH_OK();
  return S3D_OK;
}

}  // extern "C"
