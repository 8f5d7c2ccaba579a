// Point-grouping kernels (reference data/pointnet_util.py): brute-force kNN (square_distance :22-36 + argsort[:K]
// :119-120), ball query (:76-96), farthest point sampling (:53-73) and the index_points gather (:39-50).
// Integer outputs are bit-exact with the reference: distances are evaluated as ((dx*dx)+(dy*dy))+(dz*dz) in fp32 with
// explicit round-to-nearest intrinsics (no FMA contraction) and ordering is ascending (distance, index).
// No distance matrix is ever written to HBM: candidate coordinates are staged in shared memory and every query thread
// keeps its K best in registers.
#include <stdlib.h>

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

// ------------------------------------------------------------------------------------------------
// kNN, warp per query (N <= 2048 candidates, K <= 32). The thread-per-query kernel above pays its K-deep insertion
// for a whole warp whenever ANY of its 32 queries accepts a candidate, which with K = 16 is true for ~85 % of the 1024
// candidates (100 k warp instructions per 32 queries). Here lane l of a warp holds the distances of candidates
// l, l+32, ... in registers and the selection is a filter:
//   1. T = K-th smallest of the 32 per-lane minima: an upper bound of the K-th nearest distance (they are 32 distinct
//      candidates), found with one rank computation over warp shuffles;
//   2. the candidates with d <= T (about 22 for K = 16 on scattered data) are compacted into shared memory with ballots;
//   3. every survivor's rank under the (distance, index) order is counted against the other survivors; ranks < K are
//      written straight to their output slot.
// If massive ties make more than kKnnCap candidates survive, the warp falls back to K rounds of exact warp-wide
// arg-min extraction. Ordering contract unchanged: ascending (distance, index), distances bit-exact.
// ------------------------------------------------------------------------------------------------
constexpr int kKnnCap = 128;        // survivors kept per query in the fast path
constexpr int kKnnWarps = 8;        // warps per CTA
constexpr int kKnnQueriesPerWarp = 8;

__device__ __forceinline__ bool knn_less(float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); }

template <int NPL>
__global__ void __launch_bounds__(kKnnWarps * 32) knn_warp_kernel(const float* __restrict__ xyz,
                                                                 const float* __restrict__ query,
                                                                 long long* __restrict__ idx_out,
                                                                 float* __restrict__ dist_out, int N, int S, int K) {
  extern __shared__ float knn_smem[];  // [3][N] candidate coordinates | per warp: kKnnCap distances + kKnnCap indices
  float* sx = knn_smem;
  float* sy = sx + N;
  float* sz = sy + N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sd = sz + N + warp * 2 * kKnnCap;
  int* si = reinterpret_cast<int*>(sd + kKnnCap);
  const int b = blockIdx.y;
  const float* pts = xyz + (size_t)b * N * 3;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    sx[i] = pts[3 * i];
    sy[i] = pts[3 * i + 1];
    sz[i] = pts[3 * i + 2];
  }
  __syncthreads();
  const float kInf = __int_as_float(0x7f800000);
  const int q0 = (blockIdx.x * kKnnWarps + warp) * kKnnQueriesPerWarp;
  for (int qq = 0; qq < kKnnQueriesPerWarp; ++qq) {
    const int s = q0 + qq;
    if (s >= S) break;  // warp-uniform
    const float* q = query + ((size_t)b * S + s) * 3;
    const float qx = q[0], qy = q[1], qz = q[2];
    float d[NPL];
    float m = kInf;
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
      const int j = lane + 32 * t;
      d[t] = j < N ? sqdist3(qx, qy, qz, sx[j], sy[j], sz[j]) : kInf;
      m = fminf(m, d[t]);
    }
    // rank of this lane's minimum among the 32 minima (ties by lane id): the lane of rank K-1 holds the threshold
    int r = 0;
#pragma unroll
    for (int o = 0; o < 32; ++o) {
      const float mo = __shfl_sync(0xffffffffu, m, o);
      r += (mo < m || (mo == m && o < lane)) ? 1 : 0;
    }
    const unsigned holder = __ballot_sync(0xffffffffu, r == K - 1);
    const float T = __shfl_sync(0xffffffffu, m, __ffs(holder) - 1);
    // compact the candidates with d <= T
    int total = 0;
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
      const bool pass = d[t] <= T;
      const unsigned mask = __ballot_sync(0xffffffffu, pass);
      const int pos = total + __popc(mask & ((1u << lane) - 1u));
      if (pass && pos < kKnnCap) {
        sd[pos] = d[t];
        si[pos] = lane + 32 * t;
      }
      total += __popc(mask);
    }
    __syncwarp();
    long long* o = idx_out + ((size_t)b * S + s) * K;
    float* od = dist_out != nullptr ? dist_out + ((size_t)b * S + s) * K : nullptr;
    if (total <= kKnnCap) {
      for (int e = lane; e < total; e += 32) {
        const float de = sd[e];
        const int ie = si[e];
        int rank = 0;
        for (int f = 0; f < total; ++f) rank += knn_less(sd[f], si[f], de, ie) ? 1 : 0;
        if (rank < K) {
          o[rank] = (long long)ie;
          if (od != nullptr) od[rank] = de;
        }
      }
    } else {
      // exact fallback for heavy ties: K rounds of warp-wide lexicographic arg-min over the register-resident candidates
      for (int k = 0; k < K; ++k) {
        float bd = kInf;
        int bj = 0x7fffffff;
#pragma unroll
        for (int t = 0; t < NPL; ++t) {
          const int j = lane + 32 * t;
          if (j < N && knn_less(d[t], j, bd, bj)) { bd = d[t]; bj = j; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const float od2 = __shfl_xor_sync(0xffffffffu, bd, off);
          const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
          if (knn_less(od2, oj, bd, bj)) { bd = od2; bj = oj; }
        }
        if (lane == 0) {
          o[k] = (long long)bj;
          if (od != nullptr) od[k] = bd;
        }
#pragma unroll
        for (int t = 0; t < NPL; ++t)
          if (lane + 32 * t == bj) d[t] = __int_as_float(0x7fc00000);  // NaN: never "less" again (taken)
      }
    }
    __syncwarp();  // the survivor buffer is reused by the next query
  }
}

template <int NPL>
static int launch_knn_warp(const float* xyz, const float* query, long long* idx, float* dist, int B, int N, int S, int K,
                           cudaStream_t stream) {
  const size_t smem = (size_t)3 * N * sizeof(float) + (size_t)kKnnWarps * 2 * kKnnCap * sizeof(float);
  auto kern = knn_warp_kernel<NPL>;
  if (smem > 48 * 1024) S3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int per_cta = kKnnWarps * kKnnQueriesPerWarp;
  dim3 grid((S + per_cta - 1) / per_cta, B);
  kern<<<grid, kKnnWarps * 32, smem, stream>>>(xyz, query, idx, dist, N, S, K);
  S3D_LAUNCH_OK();
  return S3D_OK;
}

int knn(const float* xyz, const float* query, long long* idx, float* dist, int B, int N, int S, int K,
        cudaStream_t stream) {
  if (B <= 0 || N <= 0 || S <= 0 || K <= 0 || K > N || K > 64 || B > 65535) return S3D_ERR_BAD_SHAPE;
  if (xyz == nullptr || query == nullptr || idx == nullptr) return S3D_ERR_NULL;
  static const bool warp_path = []() { const char* v = getenv("S3D_KNN_WARP"); return v == nullptr || v[0] != '0'; }();
  if (warp_path && K <= 32 && N <= 2048) {  // warp per query, filter + rank (see knn_warp_kernel)
    if (N <= 1024) return launch_knn_warp<32>(xyz, query, idx, dist, B, N, S, K, stream);
    return launch_knn_warp<64>(xyz, query, idx, dist, B, N, S, K, stream);
  }
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
