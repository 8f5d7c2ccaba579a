/*
 * s3d_b200.h -- C ABI of libs3d_b200.so: the sm_100a kernels behind the Simple3D-Former encoder hot path.
 *
 * The reference (VITA-Group/Simple3D-Former @ a6f74c8) has no FFI/plugin interface of its own: it is pure Python and the
 * arithmetic lives in timm==0.3.2 / torch modules. Each entry point below therefore names the reference Python symbol
 * whose arithmetic it replaces (file:line in the reference tree, or the timm-0.3.2 symbol for the encoder blocks).
 * A maintainer binds these with ctypes; simple3d_former_b200/_lib.py is that binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch allocates); the library never allocates, frees or
 *     retains a pointer; work is enqueued on `stream` (a cudaStream_t passed as void*) and the call returns at once;
 *   - return value: 0 = success; negative = S3D_ERR_* argument error (nothing was launched); positive = cudaError_t;
 *   - no exceptions cross the ABI, no global state besides cached function attributes; re-entrant per stream;
 *   - "bf16" buffers hold __nv_bfloat16; "f32" buffers float; index buffers int64_t (torch.long);
 *   - row-major everywhere, leading dimensions in ELEMENTS.
 */
#ifndef S3D_B200_H_
#define S3D_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S3D_ABI_VERSION 4

enum {
  S3D_OK = 0,
  S3D_ERR_BAD_SHAPE = -1,   /* non-positive / inconsistent sizes                         */
  S3D_ERR_UNSUPPORTED = -2, /* valid but not implemented (e.g. head_dim not in 64/192/256) */
  S3D_ERR_ALIGNMENT = -3,   /* pointer or leading dimension not 16-byte aligned           */
  S3D_ERR_NULL = -4,        /* required pointer is NULL                                   */
  S3D_ERR_DRIVER = -5,      /* cuTensorMapEncodeTiled unavailable / failed                */
  S3D_ERR_WORKSPACE = -6    /* workspace too small                                        */
};

/* epilogues of s3d_gemm_bf16 */
enum { S3D_EPI_NONE = 0, S3D_EPI_GELU = 1, S3D_EPI_DGELU = 2, S3D_EPI_RELU = 3, S3D_EPI_DRELU = 4 };

int s3d_abi_version(void);
const char* s3d_error_string(int code);

/* ---------------------------------------------------------------------------------------------------------------
 * Tensor-core GEMM (tcgen05 + TMEM + TMA):  D[M,N] = epilogue(alpha * A x B^T)
 *   replaces nn.Linear inside timm-0.3.2 Attention.qkv / Attention.proj / Mlp.fc1 / Mlp.fc2,
 *   nn.MultiheadAttention in_proj/out_proj and linear1/linear2 of the group_embed layer (vit_3d_2d_pretrain.py:381),
 *   the Conv3d patchify as a GEMM (embed_layer_3d_modality.py:22-24,52-54), and all of their backward GEMMs.
 *   A: a_mn_major == 0 -> [M,K] row-major (lda >= K);  == 1 -> stored [K,M] row-major (lda >= M)
 *   B: b_mn_major == 0 -> [N,K] row-major (ldb >= K);  == 1 -> stored [K,N] row-major (ldb >= N)
 *   epilogue order: v = alpha*acc; v += bias[n]; GELU: (aux_out = bf16(v)), v = gelu_erf(v);
 *                   DGELU: v *= gelu_erf'(aux_in[m,n]); RELU: v = max(v,0); DRELU: v = aux_in[m,n] > 0 ? v : 0;
 *                   v += residual[m,n]; D = (out_fp32 ? v : bf16(v)).
 *   residual may alias D (fp32 accumulate). batch > 1 runs `batch` independent problems with element strides.
 *   force_bn / force_cluster / force_splits: 0 = heuristic; otherwise tile N (64/128/256), cluster size along M
 *   (1/2/4, TMA multicast of the B tile) and split-K factor (fp32 outputs without activation epilogue only).
 * ------------------------------------------------------------------------------------------------------------- */
int s3d_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldd,
                  int a_mn_major, int b_mn_major, int out_fp32, float alpha, const float* bias, const float* residual,
                  int64_t ldr, int epilogue, const void* aux_in, int64_t ld_aux_in, void* aux_out, int64_t ld_aux_out,
                  int batch, int64_t batch_stride_a, int64_t batch_stride_b, int64_t batch_stride_d,
                  int64_t batch_stride_r, int force_bn, int force_cluster, int force_splits, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * LayerNorm (timm Block.norm1/norm2, VisionTransformer.norm: eps 1e-6, vit_3d_2d_pretrain.py:287; post-norm layers of
 * nn.TransformerEncoderLayer: eps 1e-5). x f32 [T,D]; optional addend fuses "x + addend" (and writes it to sum_out).
 * y_bf16 / y_f32 / mean / rstd may each be NULL. Backward adds `dres` (gradient through the residual branch) and can
 * emit a bf16 copy of dx for the next GEMM; dgamma/dbeta are ACCUMULATED (caller zeroes them). dx_colsum (optional, f32
 * [D], ACCUMULATED, needs dgamma/dbeta) receives the column sums of dx: dx is the gradient of a Linear layer's output
 * (attn.proj / mlp.fc2 of the timm Block), so this is that layer's bias gradient without a pass of its own.
 * ------------------------------------------------------------------------------------------------------------- */
int s3d_layernorm_fwd(const float* x, const float* addend, float* sum_out, const float* gamma, const float* beta,
                      void* y_bf16, float* y_f32, float* mean, float* rstd, int T, int D, float eps, void* stream);
int s3d_layernorm_bwd(const void* dy, int dy_is_bf16, const float* x, const float* gamma, const float* mean,
                      const float* rstd, const float* dres, float* dx, void* dx_bf16, float* dgamma, float* dbeta,
                      float* dx_colsum, int T, int D, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Attention core (timm-0.3.2 Attention.forward: softmax(q k^T * scale) v; F.multi_head_attention_forward inside the
 * group_embed layer). q/k/v/out addressed by element strides (batch, head, row) so both [B,N,3,H,dh] and the
 * sequence-first [S,Nb,3E] layouts are read in place. head_dim in {64,192,256}. lse [B,H,N] f32 is written by fwd and
 * read by bwd; delta [B,H,N] f32 is scratch for bwd. dq/dk/dv use the q/k/v strides, dout uses the out strides.
 * Attention-probability dropout (nn.MultiheadAttention(dropout=p) inside nn.TransformerEncoderLayer, reference
 * vit_3d_2d_pretrain.py:381, active in train()): dropout_seed = DEVICE pointer to a uint32 seed (NULL or dropout_p == 0
 * = off), dropout_site = stream id; out = ((P o mask) / (1 - p)) V with the softmax normaliser taken over the full row;
 * the mask is a counter-based hash of (seed, site, (b*H + h)*N + query, key) that bwd regenerates (tcgen05 kernels:
 * head_dim 64 / 192 only).
 * ------------------------------------------------------------------------------------------------------------- */
int s3d_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int H, int N, int head_dim,
                 int64_t qkv_batch_stride, int64_t qkv_head_stride, int64_t qkv_row_stride, int64_t o_batch_stride,
                 int64_t o_head_stride, int64_t o_row_stride, float scale, const uint32_t* dropout_seed,
                 uint32_t dropout_site, float dropout_p, void* stream);
int s3d_attn_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                 float* delta, void* dq, void* dk, void* dv, int B, int H, int N, int head_dim,
                 int64_t qkv_batch_stride, int64_t qkv_head_stride, int64_t qkv_row_stride, int64_t o_batch_stride,
                 int64_t o_head_stride, int64_t o_row_stride, float scale, const uint32_t* dropout_seed,
                 uint32_t dropout_site, float dropout_p, void* workspace, int64_t workspace_bytes, void* stream);
/* Scratch (bytes, 1 KiB aligned) with which s3d_attn_bwd runs its single-score-pass form on long sequences: the dQ kernel
 * writes the bf16 matrices P o mask and dS ([B*H, N, ceil64(N)] each) once and dK / dV are batched GEMMs over them.
 * 0 = not applicable for this shape; passing workspace = NULL (or fewer bytes) selects the recomputing two-kernel form. */
int64_t s3d_attn_bwd_workspace_bytes(int B, int H, int N, int head_dim);

/* ---------------------------------------------------------------------------------------------------------------
 * Elementwise / data-movement helpers of the path
 * ------------------------------------------------------------------------------------------------------------- */
int s3d_cast_f32_to_bf16(const float* in, void* out, int64_t n, void* stream);
/* out[c, r] = bf16(in[r, c]); in is f32 or bf16 [R,C] */
int s3d_transpose_to_bf16(const void* in, int in_is_bf16, void* out, int R, int C, int64_t ld_in, int64_t ld_out,
                          void* stream);
/* out[c] (+)= sum_t in[t, c]; Linear bias gradients */
/* fp32 strided GEMM on the CUDA cores: C[M,N] (+)= alpha * sum_k A(m,k) B(k,n) [+ bias[n]] [ReLU] [zero where gate <= 0],
 * A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn], C / gate row-major. The thin fp32 layers of the point models
 * (fc1 / fc_pos_embed stem, reference models/3DViT/model.py:236-247,310-311) and the classification heads (model.py:232,
 * vit_3d_2d_pretrain.py:366): y = x W^T + b, dx = dy W and dW = dy^T x are the same call with transposed strides. Long
 * contractions with few output tiles are split and reduced with atomicAdd (C zeroed first unless accumulate). */
int s3d_sgemm_f32(const float* A, const float* B, float* C, int M, int N, int K, int64_t sam, int64_t sak, int64_t sbk,
                  int64_t sbn, int64_t ldc, float alpha, const float* bias, int relu, const float* gate, int64_t ld_gate,
                  int accumulate, void* stream);
int s3d_colsum_bf16(const void* in, float* out, int T, int C, int64_t ld, int accumulate, void* stream);
/* Conv3d(k = s = cell) operand: x [B,1,V,V,V] -> P bf16 [B*p*p*(zsum?1:p), Kpad]; zsum sums the pz patches of a
 * column first (VoxelEmbed's mean over dim 4, embed_layer_3d_modality.py:38). in_dtype: 0 = f32, 1 = uint8/bool,
 * 2 = int32 (the dtype the reference's binvox loaders yield, data/modelnet40.py:40) -- no host-side .float() needed. */
int s3d_voxel_patch_gather(const void* x, int in_dtype, void* P, int B, int V, int cell, int patch, int Kpad, int zsum,
                           void* stream);
/* torch.optim.SGD(momentum) step (train_cls.py:91, train_partseg.py:95); same conventions as s3d_adam_step. */
int s3d_sgd_momentum_step(float* param, const float* grad, float* momentum_buf, void* shadow_bf16, int64_t n, float lr,
                          float momentum, float weight_decay, int step, const int* step_device, float grad_scale,
                          void* stream);
/* torch.optim.Adam step (train_cls_voxel.py:195) over a flat f32 segment; refreshes the bf16 shadow; grad_scale folds
 * the 1/world_size of the DDP gradient average (train_cls_voxel.py:154-165). step_device (optional, int32 on the
 * device) overrides `step` so a captured CUDA graph advances the bias correction on replay. */
int s3d_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16, int64_t n,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int step, const int* step_device,
                  float grad_scale, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Point grouping (reference data/pointnet_util.py). xyz f32 [B,N,3], query f32 [B,S,3].
 *   s3d_knn          : square_distance (:22-36) + argsort()[:, :, :K] (:119-120, :233-234; models/3DViT/model.py:23-24);
 *                      idx int64 [B,S,K] ascending (distance, index); dist (optional) f32 [B,S,K].
 *   s3d_ball_query   : query_ball_point (:76-96); radius_sq = float32(radius**2); idx int64 [B,S,nsample].
 *   s3d_fps          : farthest_point_sample (:53-73) with the random start made explicit: start int64 [B].
 *   s3d_gather_rows  : index_points (:39-50): out[b,m,:] = points[b, idx[b,m], :], points f32 [B,N,C], idx int64 [B,M].
 *   s3d_scatter_add_rows: its backward (grad_points zeroed, then accumulated).
 * ------------------------------------------------------------------------------------------------------------- */
/* Element-wise dropout of the group_embed layer (dropout1 / dropout / dropout2 of nn.TransformerEncoderLayer), same
 * counter-based mask keyed by (device seed, site, row, column); rows x cols row-major, cols % 4 == 0:
 *   s3d_dropout_add_f32 : out = residual + mask o x / (1 - p)   (residual may be NULL)
 *   s3d_dropout_bf16    : out = mask o x / (1 - p)              (bf16, in place allowed) */
int s3d_dropout_add_f32(const float* x, const float* residual, float* out, int64_t rows, int cols,
                        const uint32_t* seed, uint32_t site, float p, void* stream);
int s3d_dropout_bf16(const void* x, void* out, int64_t rows, int cols, const uint32_t* seed, uint32_t site, float p,
                     void* stream);
int s3d_knn(const float* xyz, const float* query, int64_t* idx, float* dist, int B, int N, int S, int K, void* stream);
int s3d_ball_query(const float* xyz, const float* query, int64_t* idx, int B, int N, int S, float radius_sq,
                   int nsample, void* stream);
int s3d_fps(const float* xyz, const int64_t* start, int64_t* idx, int B, int N, int npoint, void* stream);
int s3d_gather_rows(const float* points, const int64_t* idx, float* out, int B, int N, int M, int C, void* stream);
int s3d_scatter_add_rows(const float* grad_out, const int64_t* idx, float* grad_points, int B, int N, int M, int C,
                         void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Set abstraction / feature propagation of the point tokenizer (reference data/pointnet_util.py:99-138 sample_and_group,
 * :220-244 PointNetSetAbstraction.forward = [Conv2d 1x1 -> BatchNorm2d -> ReLU] x 2 -> max over the K neighbours,
 * :381-420 PointNetFeaturePropagation.forward = 3-NN inverse-distance interpolation; models/3DViT/model.py:33-72
 * TransitionDown / TransitionUp = Linear -> BatchNorm1d -> ReLU). The grouped tensor [B,S,K,3+Cf] is never built:
 *   z1[b,s,k,:] = uf[b, idx[b,s,k], :] + w1[:, 0:3] (xyz[b, idx[b,s,k]] - cxyz[b,s])        (layer 1, fp32)
 * with uf = f W1[:, 3:]^T + b1 a per-point GEMM (s3d_gemm_bf16). Rows r = (b*S + s)*K + k, groups g = b*S + s.
 * BatchNorm statistics go through `partials` f32 [P, 2, C] (one slot per CTA row-slice, P chosen by the caller) and
 * a fixed-order fp64 finalize, so forward results are bitwise reproducible.
 *   s3d_sa_group_fwd_stats   : partials <- sum z1, sum z1^2
 *   s3d_sa_group_fwd_act     : a1 bf16 [B*S*K, C1] = relu(scale * z1 + shift)                   (layer-2 GEMM operand)
 *   s3d_sa_group_reduce      : z2 f32 [G*K, C] -> zmax/zmin f32 [G,C], kmax/kmin u8 [G,C] (first k on ties), partials
 *   s3d_sa_pool_select       : out = relu(scale * zsel + shift), zsel = scale >= 0 ? zmax : zmin, ksel likewise
 *                              (= max_k relu(bn(z2_k)), :241-242)
 *   s3d_sa_dz2_expand        : dz2 bf16 [G*K, C] = scale * (dy - m1 - zhat * m2); dy = dout[g,c] on row k == ksel[g,c]
 *                              where out > 0, else 0 (BatchNorm2d backward of the pooled layer)
 *   s3d_sa_group_bwd_stats   : partials <- sum dy1, sum dy1 * zhat1; dy1 = da1 where a1 > 0
 *   s3d_sa_group_bwd_scatter : dz1 = scale * (dy1 - m1 - zhat1 * m2); duf[b, idx] += dz1 (duf zeroed by the caller);
 *                              dwx_partials f32 [P, 3, C1] <- sum dz1 * (xyz[idx] - cxyz)       (gradient of w1[:, 0:3])
 *   s3d_bn_rows_stats / s3d_bn_rows_bwd_stats : the same statistics over a plain f32 [R, C] matrix (BatchNorm1d)
 *   s3d_bn_finalize_fwd      : mean, rstd, scale = gamma * rstd, shift = beta - mean * scale; running statistics
 *                              updated in place with the unbiased variance (NULL = skip), as nn.BatchNorm*d in train()
 *   s3d_bn_finalize_bwd      : dgamma (+)= sum dy * zhat, dbeta (+)= sum dy; m1, m2 = those / count (0 when !training)
 *   s3d_bn_relu_apply        : y = relu(scale * z + shift) -> f32 and/or bf16
 *   s3d_bn_relu_bwd_apply    : dz bf16 = scale * (dy - m1 - zhat * m2), dy = dout where scale * z + shift > 0
 *   s3d_split_bf16x3         : x f32 [R,K] (row stride ldx) -> bf16 [R,3K] = [hi|lo|hi] (activations) or [hi|hi|lo]
 *                              (weights), hi = bf16(x), lo = bf16(x - hi): one GEMM over 3K gives fp32-grade products
 *                              for the small per-point Linear layers in front of a BatchNorm
 *   s3d_three_nn_interp_fwd  : out[b,n,:] = sum_j w_j feats[b, idx[b,n,j], :] (+ addend), w_j = (1/(d_j+1e-8)) / sum
 *                              (:401-408); idx int64 [B,N,3], dist f32 [B,N,3] from s3d_knn(K=3)
 *   s3d_three_nn_interp_bwd  : dfeats[b, idx, :] += w_j dout[b,n,:] (dfeats zeroed here)
 * ------------------------------------------------------------------------------------------------------------- */
int s3d_sa_group_fwd_stats(const float* uf, const float* xyz, const float* cxyz, const int64_t* idx, const float* w1,
                           int ldw, int B, int N, int S, int K, int C1, float* partials, int P, void* stream);
int s3d_sa_group_fwd_act(const float* uf, const float* xyz, const float* cxyz, const int64_t* idx, const float* w1,
                         int ldw, int B, int N, int S, int K, int C1, const float* scale, const float* shift,
                         void* a1_bf16, int P, void* stream);
int s3d_sa_group_bwd_stats(const float* uf, const float* xyz, const float* cxyz, const int64_t* idx, const float* w1,
                           int ldw, int B, int N, int S, int K, int C1, const float* mean, const float* rstd,
                           const void* a1_bf16, const void* da1_bf16, float* partials, int P, void* stream);
int s3d_sa_group_bwd_scatter(const float* uf, const float* xyz, const float* cxyz, const int64_t* idx, const float* w1,
                             int ldw, int B, int N, int S, int K, int C1, const float* scale, const float* mean,
                             const float* rstd, const float* m1, const float* m2, const void* a1_bf16,
                             const void* da1_bf16, float* duf, float* dwx_partials, int P, void* stream);
int s3d_sa_group_reduce(const float* z2, int64_t G, int K, int C, float* zmax, float* zmin, uint8_t* kmax,
                        uint8_t* kmin, float* partials, int P, void* stream);
int s3d_sa_pool_select(const float* zmax, const float* zmin, const uint8_t* kmax, const uint8_t* kmin,
                       const float* scale, const float* shift, float* out, float* zsel, uint8_t* ksel, int64_t G, int C,
                       void* stream);
int s3d_sa_dz2_expand(const float* z2, const float* dout, const float* zsel, const uint8_t* ksel, const float* scale,
                      const float* shift, const float* mean, const float* rstd, const float* m1, const float* m2,
                      void* dz2_bf16, int64_t G, int K, int C, int P, void* stream);
int s3d_bn_rows_stats(const float* z, int64_t R, int C, float* partials, int P, void* stream);
int s3d_bn_rows_bwd_stats(const float* dout, const float* z, const float* scale, const float* shift, const float* mean,
                          const float* rstd, int64_t R, int C, float* partials, int P, void* stream);
int s3d_bn_finalize_fwd(const float* partials, int P, int C, double count, const float* gamma, const float* beta,
                        float eps, float momentum, float* running_mean, float* running_var, float* mean, float* rstd,
                        float* scale, float* shift, void* stream);
int s3d_bn_finalize_bwd(const float* partials, int P, int C, double count, int training, float* m1, float* m2,
                        float* dgamma, float* dbeta, int accumulate, void* stream);
int s3d_bn_relu_apply(const float* z, const float* scale, const float* shift, float* y_f32, void* y_bf16, int64_t R,
                      int C, void* stream);
int s3d_bn_relu_bwd_apply(const float* dout, const float* z, const float* scale, const float* shift, const float* mean,
                          const float* rstd, const float* m1, const float* m2, void* dz_bf16, int64_t R, int C,
                          void* stream);
int s3d_split_bf16x3(const float* x, void* out_bf16, int64_t R, int K, int64_t ldx, int weight_layout, void* stream);
int s3d_three_nn_interp_fwd(const float* feats, const int64_t* idx, const float* dist, const float* addend, float* out,
                            int B, int S, int N, int C, void* stream);
int s3d_three_nn_interp_bwd(const float* dout, const int64_t* idx, const float* dist, float* dfeats, int B, int S, int N,
                            int C, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * binvox payload -> dense occupancy grid (reference utils/binvox_rw.py:117-151 read_as_3d_array, as called by
 * data/modelnet40.py:35-45 / data/shapenet_v2.py:34-44). `payload` holds the (value, count) byte pairs of B models
 * back to back (everything after the "data\n" header line); offsets int64 [B+1] are byte offsets of each model's
 * pairs; run_offsets int64 [B+1] the same in runs (= offsets / 2 when the models are packed pair-aligned).
 *   s3d_binvox_scan   : run_end[r] = inclusive prefix sum of the counts of model b (uint32, workspace of total-runs
 *                       entries), totals[b] = number of voxels the stream encodes (the caller checks it equals V^3)
 *   s3d_binvox_expand : out [B,1,V,V,V] (out_dtype 0 = f32, 1 = u8, 2 = i32), voxel != 0 -> 1; the stream is x-z-y
 *                       ordered, fix_coords = 1 writes x-y-z like read_as_3d_array(fix_coords=True) (:143-146)
 * ------------------------------------------------------------------------------------------------------------- */
int s3d_binvox_scan(const uint8_t* payload, const int64_t* offsets, uint32_t* run_end, const int64_t* run_offsets,
                    int64_t* totals, int B, void* stream);
int s3d_binvox_expand(const uint8_t* payload, const int64_t* offsets, const uint32_t* run_end,
                      const int64_t* run_offsets, void* out, int out_dtype, int B, int V, int fix_coords, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* S3D_B200_H_ */
