// fp32 strided GEMM on the CUDA cores for the layers of the path that are too small / too thin for a tensor-core tile and
// must stay fp32: the point models' stem (fc1 = Linear(d_points, q) -> ReLU -> Linear(q, q), fc_pos_embed = Linear(3, q)
// -> ReLU -> Linear(q, q), reference models/3DViT/model.py:236-247, q = embed_dim / 4 = 48 for deit_tiny) and the
// classification heads (Linear(q, n_classes) / Linear(embed_dim, n_classes), model.py:232, vit_3d_2d_pretrain.py:366).
// Forward y = x W^T + b, input gradient dx = dy W and weight gradient dW = dy^T x are the same kernel: both operands are
// addressed through (row, column) strides, so transposed views cost nothing.
//
//   C[M,N] (+)= alpha * sum_k A(m,k) B(k,n)  [+ bias[n]]  [ReLU]  [zeroed where gate[m,n] <= 0]
//
// 64 x 64 tile per CTA, 16-deep k slices through shared memory, 4 x 4 outputs per thread. Long contractions with few
// output tiles (weight / bias gradients: K = all points of the batch) are split over grid.z and reduced with atomicAdd.
#include "kernels.h"

namespace s3d {
namespace {

constexpr int kBM = 64, kBN = 64, kBK = 16;

__global__ void __launch_bounds__(256)
sgemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int K,
                 long long sam, long long sak, long long sbk, long long sbn, long long ldc, float alpha,
                 const float* __restrict__ bias, int relu, const float* __restrict__ gate, long long ld_gate, int accumulate,
                 int k_per_split, int atomic_out) {
  pdl_prologue();
  __shared__ float As[kBK][kBM + 1];
  __shared__ __align__(16) float Bs[kBK][kBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = (sak == 1), b_nfast = (sbn == 1);
  for (int k0 = k_begin; k0 < k_end; k0 += kBK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int kk, mm;
      if (a_kfast) { kk = e % kBK; mm = e / kBK; } else { mm = e % kBM; kk = e / kBM; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < k_end) ? __ldg(A + (long long)m * sam + (long long)k * sak) : 0.f;
      int kb, nn;
      if (b_nfast) { nn = e % kBN; kb = e / kBN; } else { kb = e % kBK; nn = e / kBK; }
      const int n = n0 + nn, k2 = k0 + kb;
      Bs[kb][nn] = (n < N && k2 < k_end) ? __ldg(B + (long long)k2 * sbk + (long long)n * sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] * alpha;
      if (bias != nullptr && blockIdx.z == 0) v += __ldg(bias + n);
      float* c = C + (long long)m * ldc + n;
      if (atomic_out) {
        atomicAdd(c, v);
      } else {
        if (relu) v = fmaxf(v, 0.f);
        if (gate != nullptr && !(__ldg(gate + (long long)m * ld_gate + n) > 0.f)) v = 0.f;
        *c = accumulate ? *c + v : v;
      }
    }
  }
}

}  // namespace

int sgemm_f32(const float* A, const float* B, float* C, int M, int N, int K, long long sam, long long sak, long long sbk,
              long long sbn, long long ldc, float alpha, const float* bias, int relu, const float* gate, long long ld_gate,
              int accumulate, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return S3D_ERR_BAD_SHAPE;
  if (A == nullptr || B == nullptr || C == nullptr) return S3D_ERR_NULL;
  const int tm = (M + kBM - 1) / kBM, tn = (N + kBN - 1) / kBN;
  if (tm > 65535) return S3D_ERR_BAD_SHAPE;
  int splits = 1;
  const long long tiles = (long long)tm * tn;
  if (!relu && gate == nullptr && tiles < num_sms() && K >= 4096) {
    splits = (int)((4LL * num_sms() + tiles - 1) / tiles);  // measured on [48, 131072] x [131072, 48]: 72 -> 54 us
    const int max_splits = K / 128;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  int k_per_split = (K + splits - 1) / splits;
  k_per_split = (k_per_split + kBK - 1) / kBK * kBK;
  splits = (K + k_per_split - 1) / k_per_split;
  if (splits > 1 && !accumulate) {
    if (ldc == N) S3D_CUDA_OK(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, stream));
    else S3D_CUDA_OK(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, stream));
  }
  S3D_CUDA_OK(launch_pdl(sgemm_f32_kernel, dim3(tn, tm, splits), dim3(256), (size_t)0, stream, A, B, C, M, N, K, sam, sak,
                         sbk, sbn, ldc, alpha, bias, relu, gate, ld_gate, accumulate, k_per_split, splits > 1 ? 1 : 0));
  S3D_LAUNCH_OK();
  return S3D_OK;
}

}  // namespace s3d
