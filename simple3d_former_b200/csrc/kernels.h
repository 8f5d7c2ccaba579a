// Internal declarations shared by the translation units of libs3d_b200.so (not part of the public ABI).
#pragma once
#include "common.cuh"

namespace s3d {
// gemm_tcgen05.cu
enum : int { EPI_NONE = 0, EPI_GELU = 1, EPI_DGELU = 2, EPI_RELU = 3, EPI_DRELU = 4 };
struct GemmParams {
  int M, N, K;
  void* D;
  long long ldd;
  int out_fp32;
  const float* bias;
  const float* residual;
  long long ldr;
  const __nv_bfloat16* aux_in;
  long long ld_aux_in;
  __nv_bfloat16* aux_out;
  long long ld_aux_out;
  int epilogue;
  float alpha;
  long long batch_stride_d, batch_stride_r;
  int batched;
  // batch count and the batch-index -> tensor-map coordinate mapping (bt / batch_inner) * bmul_x + bt % batch_inner of
  // A, B and D / residual (plain batches: batch_inner = 1, bmul = 1); filled by gemm_bf16 from GemmArgs
  int batch, batch_inner, bmul_a, bmul_b, bmul_d;
  int chunk_a, chunk_b;  // MN-major operand read through a rank-4 chunk view (one TMA load per k-block); set by gemm_bf16
  int a_panel;  // A stored as 64-column panels (make_tmap_bf16_panel), addressed through rank-4 maps. 1: MN-major A, panels
                // [batch][M / 64][K][64]; 2: K-major A, panels [batch][K / 64][M][64] (the panel index is the k-block)
  // UMMA smem-descriptor byte offsets (defaults: MN-major LBO 8192 / SBO 1024, K-major LBO 16 / SBO 1024);
  // overridable through S3D_DBG_* environment variables for bring-up on new silicon.
  unsigned mn_lbo, mn_sbo, k_lbo, k_sbo;
  int splits;  // split-K factor (> 1: fp32 red.add epilogue into a zeroed / accumulating D)
};
struct GemmArgs {
  const void* A;
  const void* B;
  long long lda, ldb;
  int a_mn, b_mn;
  int batch;
  long long batch_stride_a, batch_stride_b;
  // optional two-level batches (attention: batch = B * H, batch_inner = H): operand x of batch bt sits at tensor-map
  // batch coordinate (bt / batch_inner) * bmul_x + bt % batch_inner, in units of its batch stride. 0 = plain batches.
  int batch_inner, bmul_a, bmul_b, bmul_d;
  int force_bn;
  int force_cluster;  // 0 = auto
  int force_splits;   // 0 = auto
  GemmParams p;
};
int gemm_bf16(const GemmArgs& g, cudaStream_t stream);

// norm_elementwise.cu
int layernorm_fwd(const float* x, const float* addend, float* sum_out, const float* gamma, const float* beta,
                  void* y_bf16, float* y_f32, float* mean, float* rstd, int T, int D, float eps, cudaStream_t stream);
int layernorm_bwd(const void* dy, int dy_is_bf16, const float* x, const float* gamma, const float* mean,
                  const float* rstd, const float* dres, float* dx, void* dx_bf16, float* dgamma, float* dbeta, float* dxsum, int T,
                  int D, cudaStream_t stream);
int cast_f32_to_bf16(const float* in, void* out, long long n, cudaStream_t stream);
int transpose_to_bf16(const void* in, int in_is_bf16, void* out, int R, int C, long long ld_in, long long ld_out,
                      cudaStream_t stream);
int colsum_bf16(const void* in, float* out, int T, int C, long long ld, int accumulate, cudaStream_t stream);
int dropout_add_f32(const float* x, const float* res, float* out, long long rows, int cols, const uint32_t* seed,
                    unsigned site, float p, cudaStream_t stream);
int dropout_bf16(const void* x, void* out, long long rows, int cols, const uint32_t* seed, unsigned site, float p,
                 cudaStream_t stream);
int voxel_patch_gather(const void* x, int in_dtype, void* P, int B, int V, int cell, int patch, int Kpad, int zsum,
                       cudaStream_t stream);
int sgd_momentum_step(float* p, const float* g, float* buf, void* shadow_bf16, long long n, float lr, float momentum,
                      float weight_decay, int step, const int* step_dev, float grad_scale, cudaStream_t stream);
int adam_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, long long n, float lr, float beta1,
              float beta2, float eps, float weight_decay, int step, const int* step_dev, float grad_scale,
              cudaStream_t stream);

// attention_mma.cu
struct AttnParams {
  const __nv_bfloat16 *q, *k, *v;
  const __nv_bfloat16 *o, *dout;
  __nv_bfloat16* out;
  __nv_bfloat16 *dq, *dk, *dv;
  float* lse;
  float* delta;
  long long qkv_bs, qkv_hs, qkv_rs;
  long long o_bs, o_hs, o_rs;
  int B, H, N;
  float scale;
  // attention-probability dropout (tcgen05 kernels only): device seed (nullptr = off), site id, threshold, 1 / (1 - p)
  const uint32_t* drop_seed;
  uint32_t drop_site;
  uint32_t drop_thresh14;  // keep <=> 14-bit draw >= thresh14 = round(p * 16384), see common.cuh
  float drop_scale;
  // backward only: caller-owned scratch for the single-score-pass form (attn_bwd_workspace_bytes; nullptr = two-kernel form)
  void* workspace;
  long long workspace_bytes;
};
int attn_fwd(const AttnParams& p, int DH, cudaStream_t stream);
int attn_bwd(const AttnParams& p, int DH, cudaStream_t stream);
// attention_tcgen05.cu (long sequences / dropout, dh in {48, 64, 96, 192}); S3D_ERR_UNSUPPORTED -> the layout is not a
// slice of one 2-D qkv buffer and the caller uses the mma.sync kernels
bool attn_tc_supported(int DH);      // forward and backward
bool attn_tc_fwd_supported(int DH);  // forward only: additionally head_dim 256 (all A operands fit tensor memory)
int attn_fwd_tc(const AttnParams& p, int DH, cudaStream_t stream);
int attn_bwd_tc(const AttnParams& p, int DH, cudaStream_t stream);
long long attn_bwd_workspace_bytes(int B, int H, int N, int DH);

// linear_f32.cu
int sgemm_f32(const float* A, const float* B, float* C, int M, int N, int K, long long sam, long long sak, long long sbk,
              long long sbn, long long ldc, float alpha, const float* bias, int relu, const float* gate, long long ld_gate,
              int accumulate, cudaStream_t stream);

// pointops.cu
int knn(const float* xyz, const float* query, long long* idx, float* dist, int B, int N, int S, int K,
        cudaStream_t stream);
int ball_query(const float* xyz, const float* query, long long* idx, int B, int N, int S, float radius_sq, int nsample,
               cudaStream_t stream);
int fps(const float* xyz, const long long* start, long long* out, int B, int N, int npoint, cudaStream_t stream);
int gather_rows(const float* points, const long long* idx, float* out, int B, int N, int M, int C,
                cudaStream_t stream);
int scatter_add_rows(const float* grad_out, const long long* idx, float* grad_points, int B, int N, int M, int C,
                     cudaStream_t stream);
}  // namespace s3d
