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
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (long long base = 0; base < nruns; base += kScanThreads) {
    const long long r = base + threadIdx.x;
    unsigned v = r < nruns ? (unsigned)pairs[2 * r + 1] : 0u;
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
    const unsigned carry = carry_s;
    const unsigned incl = v + (warp > 0 ? warp_sums[warp - 1] : 0u) + carry;
    if (r < nruns) out[r] = incl;
    __syncthreads();
    if (threadIdx.x == kScanThreads - 1) carry_s = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[b] = (long long)carry_s;
}

template <typename TOut>
__global__ void __launch_bounds__(256) binvox_expand_kernel(const uint8_t* __restrict__ payload,
                                                            const long long* __restrict__ offsets,
                                                            const unsigned* __restrict__ run_end,
                                                            const long long* __restrict__ run_offsets,
                                                            TOut* __restrict__ out, int V, int fix_coords) {
  extern __shared__ uint8_t plane[];  // [V][V] in stream order (z-major, y fastest)
  const int b = blockIdx.y, x = blockIdx.x;
  const uint8_t* pairs = payload + offsets[b];
  const unsigned* ends = run_end + run_offsets[b];
  const int nruns = (int)((offsets[b + 1] - offsets[b]) / 2);
  const unsigned slab = (unsigned)V * (unsigned)V;
  const unsigned lo = (unsigned)x * slab, hi = lo + slab;
  for (int i = threadIdx.x; i < (int)slab; i += blockDim.x) plane[i] = 0;  // voxels past a short stream read as empty
  // first run whose end is > lo
  int first = 0;
  {
    int l = 0, h = nruns;
    while (l < h) {
      const int m = (l + h) >> 1;
      if (ends[m] > lo) h = m; else l = m + 1;
    }
    first = l;
  }
  __syncthreads();
  for (int r = first + threadIdx.x; r < nruns; r += blockDim.x) {
    const unsigned e = ends[r];
    const unsigned cnt = pairs[2 * r + 1];
    const unsigned s = e - cnt;
    if (s >= hi) break;  // runs are ordered: everything this thread would see later is past the slab too
    if (pairs[2 * r] != 0) {
      const unsigned a = s > lo ? s : lo, z = e < hi ? e : hi;
      for (unsigned i = a; i < z; ++i) plane[i - lo] = 1;
    }
  }
  __syncthreads();
  // write out: out[b][x][j][k]; fix_coords: (j, k) = (y, z) <- plane[z][y]; otherwise the stream order (z, y) is kept
  TOut* o = out + ((size_t)b * V + x) * slab;
  for (int i = threadIdx.x; i < (int)slab; i += blockDim.x) {
    const int j = i / V, k = i % V;
    const uint8_t v = fix_coords ? plane[k * V + j] : plane[i];
    o[i] = (TOut)v;
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
  const size_t smem = (size_t)V * V;
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
