// extern "C" surface of libs3d_b200.so (declared in include/s3d_b200.h). Plain pointers and sizes only.
#include "../../include/s3d_b200.h"

#include "kernels.h"



static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int s3d_abi_version(void) { return S3D_ABI_VERSION; }

const char* s3d_error_string(int code) {
  switch (code) {
    case S3D_OK: return "ok";
    case S3D_ERR_BAD_SHAPE: return "bad shape";
    case S3D_ERR_UNSUPPORTED: return "unsupported configuration";
    case S3D_ERR_ALIGNMENT: return "pointer or leading dimension not 16-byte aligned";
    case S3D_ERR_NULL: return "required pointer is NULL";
    case S3D_ERR_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    case S3D_ERR_WORKSPACE: return "workspace too small";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown error";
}

int s3d_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldd,
                  int a_mn_major, int b_mn_major, int out_fp32, float alpha, const float* bias, const float* residual,
                  int64_t ldr, int epilogue, const void* aux_in, int64_t ld_aux_in, void* aux_out, int64_t ld_aux_out,
                  int batch, int64_t batch_stride_a, int64_t batch_stride_b, int64_t batch_stride_d,
                  int64_t batch_stride_r, int force_bn, int force_cluster, int force_splits, void* stream) {
  if (epilogue < S3D_EPI_NONE || epilogue > S3D_EPI_DRELU) return S3D_ERR_UNSUPPORTED;
  if (batch < 1 || batch > 65535) return S3D_ERR_BAD_SHAPE;
  s3d::GemmArgs g{};
  g.A = A;
  g.B = B;
  g.lda = lda;
  g.ldb = ldb;
  g.a_mn = a_mn_major ? 1 : 0;
  g.b_mn = b_mn_major ? 1 : 0;
  g.batch = batch;
  g.batch_stride_a = batch_stride_a;
  g.batch_stride_b = batch_stride_b;
  g.force_bn = force_bn;
  g.force_cluster = force_cluster;
  g.force_splits = force_splits;
  g.p.M = M;
  g.p.N = N;
  g.p.K = K;
  g.p.D = D;
  g.p.ldd = ldd;
  g.p.out_fp32 = out_fp32 ? 1 : 0;
  g.p.bias = bias;
  g.p.residual = residual;
  g.p.ldr = ldr;
  g.p.aux_in = reinterpret_cast<const __nv_bfloat16*>(aux_in);
  g.p.ld_aux_in = ld_aux_in;
  g.p.aux_out = reinterpret_cast<__nv_bfloat16*>(aux_out);
  g.p.ld_aux_out = ld_aux_out;
  g.p.epilogue = epilogue;
  g.p.alpha = alpha;
  g.p.batch_stride_d = batch_stride_d;
  g.p.batch_stride_r = batch_stride_r;
  g.p.batched = batch > 1 ? 1 : 0;
  return s3d::gemm_bf16(g, as_stream(stream));
}

int s3d_sgemm_f32(const float* A, const float* B, float* C, int M, int N, int K, int64_t sam, int64_t sak, int64_t sbk,
                  int64_t sbn, int64_t ldc, float alpha, const float* bias, int relu, const float* gate, int64_t ld_gate,
                  int accumulate, void* stream) {
  return s3d::sgemm_f32(A, B, C, M, N, K, sam, sak, sbk, sbn, ldc, alpha, bias, relu, gate, ld_gate, accumulate,
                        as_stream(stream));
}

int s3d_layernorm_fwd(const float* x, const float* addend, float* sum_out, const float* gamma, const float* beta,
                      void* y_bf16, float* y_f32, float* mean, float* rstd, int T, int D, float eps, void* stream) {
  return s3d::layernorm_fwd(x, addend, sum_out, gamma, beta, y_bf16, y_f32, mean, rstd, T, D, eps, as_stream(stream));
}

int s3d_layernorm_bwd(const void* dy, int dy_is_bf16, const float* x, const float* gamma, const float* mean,
                      const float* rstd, const float* dres, float* dx, void* dx_bf16, float* dgamma, float* dbeta,
                      float* dx_colsum, int T, int D, void* stream) {
  return s3d::layernorm_bwd(dy, dy_is_bf16, x, gamma, mean, rstd, dres, dx, dx_bf16, dgamma, dbeta, dx_colsum, T, D,
                            as_stream(stream));
}

static s3d::AttnParams make_attn(const void* q, const void* k, const void* v, int B, int H, int N, int64_t qbs,
                                 int64_t qhs, int64_t qrs, int64_t obs, int64_t ohs, int64_t ors, float scale) {
  s3d::AttnParams p{};
  p.q = reinterpret_cast<const __nv_bfloat16*>(q);
  p.k = reinterpret_cast<const __nv_bfloat16*>(k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(v);
  p.B = B;
  p.H = H;
  p.N = N;
  p.qkv_bs = qbs;
  p.qkv_hs = qhs;
  p.qkv_rs = qrs;
  p.o_bs = obs;
  p.o_hs = ohs;
  p.o_rs = ors;
  p.scale = scale;
  return p;
}

static int set_dropout(s3d::AttnParams& p, const uint32_t* seed, uint32_t site, float prob) {
  if (seed == nullptr || prob <= 0.f) return 0;
  if (prob >= 1.f) return s3d::S3D_ERR_BAD_SHAPE;
  p.drop_seed = seed;
  p.drop_site = site;
  p.drop_thresh14 = s3d::drop_thresh14(prob);
  p.drop_scale = s3d::drop_keep_scale(p.drop_thresh14);
  return 0;
}

int s3d_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int H, int N, int head_dim,
                 int64_t qkv_batch_stride, int64_t qkv_head_stride, int64_t qkv_row_stride, int64_t o_batch_stride,
                 int64_t o_head_stride, int64_t o_row_stride, float scale, const uint32_t* dropout_seed,
                 uint32_t dropout_site, float dropout_p, void* stream) {
  s3d::AttnParams p = make_attn(q, k, v, B, H, N, qkv_batch_stride, qkv_head_stride, qkv_row_stride, o_batch_stride,
                                o_head_stride, o_row_stride, scale);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  if (int rc = set_dropout(p, dropout_seed, dropout_site, dropout_p)) return rc;
  return s3d::attn_fwd(p, head_dim, as_stream(stream));
}

int64_t s3d_attn_bwd_workspace_bytes(int B, int H, int N, int head_dim) {
  return s3d::attn_bwd_workspace_bytes(B, H, N, head_dim);
}

int s3d_attn_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                 float* delta, void* dq, void* dk, void* dv, int B, int H, int N, int head_dim,
                 int64_t qkv_batch_stride, int64_t qkv_head_stride, int64_t qkv_row_stride, int64_t o_batch_stride,
                 int64_t o_head_stride, int64_t o_row_stride, float scale, const uint32_t* dropout_seed,
                 uint32_t dropout_site, float dropout_p, void* workspace, int64_t workspace_bytes, void* stream) {
  s3d::AttnParams p = make_attn(q, k, v, B, H, N, qkv_batch_stride, qkv_head_stride, qkv_row_stride, o_batch_stride,
                                o_head_stride, o_row_stride, scale);
  if (int rc = set_dropout(p, dropout_seed, dropout_site, dropout_p)) return rc;
  if (workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 1023) != 0) return S3D_ERR_ALIGNMENT;
  p.workspace = workspace;
  p.workspace_bytes = workspace_bytes;
  p.o = reinterpret_cast<const __nv_bfloat16*>(out);
  p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  p.lse = const_cast<float*>(lse);
  p.delta = delta;
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq);
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  return s3d::attn_bwd(p, head_dim, as_stream(stream));
}

int s3d_cast_f32_to_bf16(const float* in, void* out, int64_t n, void* stream) {
  return s3d::cast_f32_to_bf16(in, out, n, as_stream(stream));
}
int s3d_transpose_to_bf16(const void* in, int in_is_bf16, void* out, int R, int C, int64_t ld_in, int64_t ld_out,
                          void* stream) {
  return s3d::transpose_to_bf16(in, in_is_bf16, out, R, C, ld_in, ld_out, as_stream(stream));
}
int s3d_colsum_bf16(const void* in, float* out, int T, int C, int64_t ld, int accumulate, void* stream) {
  return s3d::colsum_bf16(in, out, T, C, ld, accumulate, as_stream(stream));
}
int s3d_voxel_patch_gather(const void* x, int in_dtype, void* P, int B, int V, int cell, int patch, int Kpad, int zsum,
                           void* stream) {
  return s3d::voxel_patch_gather(x, in_dtype, P, B, V, cell, patch, Kpad, zsum, as_stream(stream));
}
int s3d_sgd_momentum_step(float* param, const float* grad, float* momentum_buf, void* shadow_bf16, int64_t n, float lr,
                          float momentum, float weight_decay, int step, const int* step_device, float grad_scale,
                          void* stream) {
  return s3d::sgd_momentum_step(param, grad, momentum_buf, shadow_bf16, n, lr, momentum, weight_decay, step, step_device,
                                grad_scale, as_stream(stream));
}
int s3d_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16, int64_t n,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int step, const int* step_device,
                  float grad_scale, void* stream) {
  return s3d::adam_step(param, grad, exp_avg, exp_avg_sq, shadow_bf16, n, lr, beta1, beta2, eps, weight_decay, step,
                        step_device, grad_scale, as_stream(stream));
}

int s3d_knn(const float* xyz, const float* query, int64_t* idx, float* dist, int B, int N, int S, int K, void* stream) {
  return s3d::knn(xyz, query, reinterpret_cast<long long*>(idx), dist, B, N, S, K, as_stream(stream));
}
int s3d_ball_query(const float* xyz, const float* query, int64_t* idx, int B, int N, int S, float radius_sq,
                   int nsample, void* stream) {
  return s3d::ball_query(xyz, query, reinterpret_cast<long long*>(idx), B, N, S, radius_sq, nsample,
                         as_stream(stream));
}
int s3d_fps(const float* xyz, const int64_t* start, int64_t* idx, int B, int N, int npoint, void* stream) {
  return s3d::fps(xyz, reinterpret_cast<const long long*>(start), reinterpret_cast<long long*>(idx), B, N, npoint,
                  as_stream(stream));
}
int s3d_dropout_add_f32(const float* x, const float* residual, float* out, int64_t rows, int cols,
                        const uint32_t* seed, uint32_t site, float p, void* stream) {
  return s3d::dropout_add_f32(x, residual, out, rows, cols, seed, site, p, as_stream(stream));
}
int s3d_dropout_bf16(const void* x, void* out, int64_t rows, int cols, const uint32_t* seed, uint32_t site, float p,
                     void* stream) {
  return s3d::dropout_bf16(x, out, rows, cols, seed, site, p, as_stream(stream));
}
int s3d_gather_rows(const float* points, const int64_t* idx, float* out, int B, int N, int M, int C, void* stream) {
  return s3d::gather_rows(points, reinterpret_cast<const long long*>(idx), out, B, N, M, C, as_stream(stream));
}
int s3d_scatter_add_rows(const float* grad_out, const int64_t* idx, float* grad_points, int B, int N, int M, int C,
                         void* stream) {
  return s3d::scatter_add_rows(grad_out, reinterpret_cast<const long long*>(idx), grad_points, B, N, M, C,
                               as_stream(stream));
}

}  // extern "C"
