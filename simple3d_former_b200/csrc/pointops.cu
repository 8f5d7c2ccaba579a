// Point-grouping kernels (reference data/pointnet_util.py): brute-force kNN (square_distance :22-36 + argsort[:K]
// :119-120), ball query (:76-96), farthest point sampling (:53-73) and the index_points gather (:39-50).
// Integer outputs are bit-exact with the reference: distances are evaluated as ((dx*dx)+(dy*dy))+(dz*dz) in fp32 with
// explicit round-to-nearest intrinsics (no FMA contraction) and ordering is ascending (distance, index).
// No distance matrix is ever written to HBM: candidate coordinates are staged in shared memory and every query thread
// keeps its K best in registers.
#include "kernels.h"

namespace s3d {

__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

constexpr int kKnnTile = 2048;  // candidate points staged per shared-memory pass (24 KB)

// ------------------------------------------------------------------------------------------------
// kNN: one thread per query, candidates streamed through shared memory in ascending index order.
// A strict '<' against the current worst keeps the earlier index on exact distance ties (stable argsort order).
// ------------------------------------------------------------------------------------------------
template <int KMAX>
__global__ void __launch_bounds__(128) knn_kernel(const float* __restrict__ xyz, const float* __restrict__ query,
                                                 long long* __restrict__ idx_out, float* __restrict__ dist_out, int N,
                                                 int S, int K) {
  __shared__ float sx[kKnnTile], sy[kKnnTile], sz[kKnnTile];
  const int b = blockIdx.y;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = s < S;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (active) {
    const float* q = query + ((size_t)b * S + s) * 3;
    qx = q[0]; qy = q[1]; qz = q[2];
  }
  float bd[KMAX];
  int bi[KMAX];
#pragma unroll
  for (int t = 0; t < KMAX; ++t) { bd[t] = __int_as_float(0x7f800000); bi[t] = 0x7fffffff; }

  const float* pts = xyz + (size_t)b * N * 3;
  for (int base = 0; base < N; base += kKnnTile) {
    const int cnt = min(kKnnTile, N - base);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const float* p = pts + (size_t)(base + i) * 3;
      sx[i] = p[0]; sy[i] = p[1]; sz[i] = p[2];
    }
    __syncthreads();
    if (active) {
      for (int j = 0; j < cnt; ++j) {
        const float d = sqdist3(qx, qy, qz, sx[j], sy[j], sz[j]);
        if (d < bd[KMAX - 1]) {
          float cd = d;
          int ci = base + j;
          bool placed = false;  // once the new entry is placed, every later slot shifts down by one (ties included)
#pragma unroll
          for (int t = 0; t < KMAX; ++t) {
            if (placed || cd < bd[t]) {
              const float td = bd[t]; const int ti = bi[t];
              bd[t] = cd; bi[t] = ci;
              cd = td; ci = ti;
              placed = true;
            }
          }
        }
      }
    }
  }
  if (active) {
    long long* o = idx_out + ((size_t)b * S + s) * K;
#pragma unroll
    for (int t = 0; t < KMAX; ++t)
      if (t < K) o[t] = (long long)bi[t];
    if (dist_out != nullptr) {
      float* od = dist_out + ((size_t)b * S + s) * K;
#pragma unroll
      for (int t = 0; t < KMAX; ++t)
        if (t < K) od[t] = bd[t];
    }
  }
}

int knn(const float* xyz, const float* query, long long* idx, float* dist, int B, int N, int S, int K,
        cudaStream_t stream) {
  if (B <= 0 || N <= 0 || S <= 0 || K <= 0 || K > N || K > 64 || B > 65535) return S3D_ERR_BAD_SHAPE;
  if (xyz == nullptr || query == nullptr || idx == nullptr) return S3D_ERR_NULL;
  dim3 grid((S + 127) / 128, B);
  // The register list holds the KMAX best; its first K entries are the K best (buckets bound compile time).
  if (K <= 4) knn_kernel<4><<<grid, 128, 0, stream>>>(xyz, query, idx, dist, N, S, K);
  else if (K <= 16) knn_kernel<16><<<grid, 128, 0, stream>>>(xyz, query, idx, dist, N, S, K);
  else if (K <= 32) knn_kernel<32><<<grid, 128, 0, stream>>>(xyz, query, idx, dist, N, S, K);
  else knn_kernel<64><<<grid, 128, 0, stream>>>(xyz, query, idx, dist, N, S, K);
  S3D_LAUNCH_OK();
  return S3D_OK;
}

// ------------------------------------------------------------------------------------------------
// Ball query: first `nsample` candidate indices (ascending) with !(d > r2); short rows are padded with the first hit,
// rows with no hit are filled with N exactly as the reference's masked sort does.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ball_query_kernel(const float* __restrict__ xyz, const float* __restrict__ query,
                                                        long long* __restrict__ idx_out, int N, int S, float r2,
                                                        int nsample) {
  __shared__ float sx[kKnnTile], sy[kKnnTile], sz[kKnnTile];
  const int b = blockIdx.y;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = s < S;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  long long* o = nullptr;
  if (active) {
    const float* q = query + ((size_t)b * S + s) * 3;
    qx = q[0]; qy = q[1]; qz = q[2];
    o = idx_out + ((size_t)b * S + s) * nsample;
  }
  int cnt_hit = 0;
  long long first = N;
  const float* pts = xyz + (size_t)b * N * 3;
  for (int base = 0; base < N; base += kKnnTile) {
    const int cnt = min(kKnnTile, N - base);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const float* p = pts + (size_t)(base + i) * 3;
      sx[i] = p[0]; sy[i] = p[1]; sz[i] = p[2];
    }
    __syncthreads();
    if (active && cnt_hit < nsample) {
      for (int j = 0; j < cnt && cnt_hit < nsample; ++j) {
        const float d = sqdist3(qx, qy, qz, sx[j], sy[j], sz[j]);
        if (!(d > r2)) {
          if (cnt_hit == 0) first = base + j;
          o[cnt_hit++] = base + j;
        }
      }
    }
  }
  if (active)
    for (int t = cnt_hit; t < nsample; ++t) o[t] = first;
}

int ball_query(const float* xyz, const float* query, long long* idx, int B, int N, int S, float radius_sq, int nsample,
               cudaStream_t stream) {
  if (B <= 0 || N <= 0 || S <= 0 || nsample <= 0 || B > 65535) return S3D_ERR_BAD_SHAPE;
  if (xyz == nullptr || query == nullptr || idx == nullptr) return S3D_ERR_NULL;
  dim3 grid((S + 127) / 128, B);
  ball_query_kernel<<<grid, 128, 0, stream>>>(xyz, query, idx, N, S, radius_sq, nsample);
  S3D_LAUNCH_OK();
  return S3D_OK;
}

// ------------------------------------------------------------------------------------------------
// Farthest point sampling: one CTA per cloud, points and running min-distances in registers, one __syncthreads per
// iteration. argmax uses redux.sync on the fp32 bit pattern (distances are non-negative) with first-index tie-break,
// matching torch.max(dim)'s first-occurrence rule on CPU.
// ------------------------------------------------------------------------------------------------
template <int PPT>
__global__ void __launch_bounds__(1024) fps_kernel(const float* __restrict__ xyz, const long long* __restrict__ start,
                                                  long long* __restrict__ out, int N, int npoint) {
  extern __shared__ float sp[];  // [3*N] coordinates, then reduction slots
  float* sxyz = sp;
  uint32_t* red_val = reinterpret_cast<uint32_t*>(sp + 3 * (size_t)N);  // [2][32]
  int* red_idx = reinterpret_cast<int*>(red_val + 64);                  // [2][32]
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const float* pts = xyz + (size_t)b * N * 3;
  for (int i = tid; i < 3 * N; i += blockDim.x) sxyz[i] = pts[i];
  float px[PPT], py[PPT], pz[PPT], md[PPT];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int i = tid + k * blockDim.x;
    if (i < N) { px[k] = sxyz[3 * i]; py[k] = sxyz[3 * i + 1]; pz[k] = sxyz[3 * i + 2]; }
    else { px[k] = 0.f; py[k] = 0.f; pz[k] = 0.f; }
    md[k] = 1e10f;
  }
  int far = (int)start[b];
  long long* o = out + (size_t)b * npoint;
  for (int it = 0; it < npoint; ++it) {
    if (tid == 0) o[it] = far;
    const float cx = sxyz[3 * far], cy = sxyz[3 * far + 1], cz = sxyz[3 * far + 2];
    uint32_t best = 0;
    int besti = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const int i = tid + k * blockDim.x;
      if (i < N) {
        const float d = sqdist3(px[k], py[k], pz[k], cx, cy, cz);
        md[k] = fminf(md[k], d);
        const uint32_t u = __float_as_uint(md[k]);
        if (u > best || besti == 0x7fffffff) { best = u; besti = i; }
      }
    }
    // warp argmax (first index on ties)
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, best);
    const int widx = __reduce_min_sync(0xffffffffu, (best == wmax) ? besti : 0x7fffffff);
    const int slot = (it & 1) * 32;
    if (lane == 0) { red_val[slot + warp] = wmax; red_idx[slot + warp] = widx; }
    __syncthreads();
    const uint32_t v = (lane < nwarps) ? red_val[slot + lane] : 0u;
    const int vi = (lane < nwarps) ? red_idx[slot + lane] : 0x7fffffff;
    const uint32_t gmax = __reduce_max_sync(0xffffffffu, v);
    far = __reduce_min_sync(0xffffffffu, (v == gmax) ? vi : 0x7fffffff);
  }
}

int fps(const float* xyz, const long long* start, long long* out, int B, int N, int npoint, cudaStream_t stream) {
  if (B <= 0 || N <= 0 || npoint <= 0 || N > 8192) return S3D_ERR_BAD_SHAPE;
  if (xyz == nullptr || start == nullptr || out == nullptr) return S3D_ERR_NULL;
  const size_t shmem = 3 * (size_t)N * sizeof(float) + 128 * sizeof(uint32_t);
  int threads, ppt;
  if (N <= 1024) { threads = 256; ppt = 4; }
  else if (N <= 2048) { threads = 512; ppt = 4; }
  else if (N <= 4096) { threads = 512; ppt = 8; }
  else { threads = 1024; ppt = 8; }
  if (ppt == 4) {
    static bool set4 = false;
    if (!set4) { S3D_CUDA_OK(cudaFuncSetAttribute(fps_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); set4 = true; }
    fps_kernel<4><<<B, threads, shmem, stream>>>(xyz, start, out, N, npoint);
  } else {
    static bool set8 = false;
    if (!set8) { S3D_CUDA_OK(cudaFuncSetAttribute(fps_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); set8 = true; }
    fps_kernel<8><<<B, threads, shmem, stream>>>(xyz, start, out, N, npoint);
  }
  S3D_LAUNCH_OK();
  return S3D_OK;
}

// ------------------------------------------------------------------------------------------------
// index_points: out[b, m, :] = points[b, idx[b, m], :]   (fp32 rows of C channels), one warp per output row.
// Backward: scatter-add of the row gradients (atomicAdd; rows may repeat).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ points,
                                                         const long long* __restrict__ idx, float* __restrict__ out,
                                                         long long rows_total, int M, int N, int C) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows_total; r += nwarps) {
    const long long b = r / M;
    const long long src = idx[r];
    const float* s = points + ((size_t)b * N + (size_t)src) * C;
    float* d = out + (size_t)r * C;
    for (int c = lane; c < C; c += 32) d[c] = s[c];
  }
}

__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const float* __restrict__ grad_out,
                                                              const long long* __restrict__ idx,
                                                              float* __restrict__ grad_points, long long rows_total,
                                                              int M, int N, int C) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows_total; r += nwarps) {
    const long long b = r / M;
    const long long dst = idx[r];
    float* d = grad_points + ((size_t)b * N + (size_t)dst) * C;
    const float* s = grad_out + (size_t)r * C;
    for (int c = lane; c < C; c += 32) atomicAdd(d + c, s[c]);
  }
}

int gather_rows(const float* points, const long long* idx, float* out, int B, int N, int M, int C,
                cudaStream_t stream) {
  if (B <= 0 || N <= 0 || M <= 0 || C <= 0) return S3D_ERR_BAD_SHAPE;
  if (points == nullptr || idx == nullptr || out == nullptr) return S3D_ERR_NULL;
  const long long rows = (long long)B * M;
  long long blocks = (rows + 7) / 8;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  gather_rows_kernel<<<(int)blocks, 256, 0, stream>>>(points, idx, out, rows, M, N, C);
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int scatter_add_rows(const float* grad_out, const long long* idx, float* grad_points, int B, int N, int M, int C,
                     cudaStream_t stream) {
  if (B <= 0 || N <= 0 || M <= 0 || C <= 0) return S3D_ERR_BAD_SHAPE;
  if (grad_out == nullptr || idx == nullptr || grad_points == nullptr) return S3D_ERR_NULL;
  S3D_CUDA_OK(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)B * N * C, stream));
  const long long rows = (long long)B * M;
  long long blocks = (rows + 7) / 8;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  scatter_add_rows_kernel<<<(int)blocks, 256, 0, stream>>>(grad_out, idx, grad_points, rows, M, N, C);
  S3D_LAUNCH_OK();
  return S3D_OK;
}

}  // namespace s3d
