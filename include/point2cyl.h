/*
 * point2cyl.h — C-ABI of libp2c.so: the sm_100a kernels behind Point2Cyl's forward+loss hot path.
 *
 * The reference (mikacuy/point2cyl) has no FFI of its own: its plug point is Python module names
 * (SURVEY.md section 8b).  This library is what the Python drop-in modules under
 * point2cyl_b200/dropin/ bind with ctypes; each entry point names the reference code it replaces
 * (paths relative to the upstream repo root).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless marked host;
 *   - no allocation, no ownership transfer: outputs and scratch are allocated by the caller;
 *   - no exceptions cross the boundary: 0 = ok, >0 = cudaError_t, <0 = argument check (P2C_E*);
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised;
 *   - no global state besides one-time cudaFuncSetAttribute per device; re-entrant per device, so
 *     one process per GPU is safe;
 *   - clouds are point-major float32: xyz (B,N,3), features (rows, C) with an explicit row stride
 *     `ld` in elements; indices are int64 like the reference's torch.long.
 */
#ifndef POINT2CYL_H
#define POINT2CYL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P2C_EINVAL (-1)      /* bad size / null pointer */
#define P2C_EUNSUPPORTED (-2) /* shape outside what this build's kernels cover */
#define P2C_EALIGN (-3)      /* pointer / stride alignment */

/* build checks */
int p2c_version(void);           /* 100 * major + minor */
const char* p2c_arch(void);      /* "sm_100a" */

/* Number of SMs the persistent one-CTA-per-SM tensor-core kernels (p2c_linear, p2c_wgrad) of SUBSEQUENT launches
 * from this process may occupy; 0 = all.  Returns the previous value.  point2cyl_b200.graph.PipelinedForwardLoss
 * sets it so that the coordinate-only stage of the next batch (FPS, ball query, 3-NN search: small grids) runs on a
 * second stream beside the per-point MLP layers of the current batch. */
int p2c_set_sm_budget(int sms);
/* Programmatic dependent launch (griddepcontrol) for the kernels that support it: their CTAs may be scheduled while the
 * predecessor in the stream drains and wait for its completion after their local prologue.  Default on; returns the
 * previous setting. */
int p2c_set_pdl(int on);

/* Farthest point sampling — replaces farthest_point_sample, models/pointnet_util.py:63-84, fused
 * with the gather of the sampled centres (index_points, :43-60, as used at :124).
 * start[b] is the first centroid (the reference draws it on the CPU generator, :75).
 * Bit-exact contract: dist = (dx*dx + dy*dy) + dz*dz with every op rounded (no FMA contraction),
 * running distance init 1e10, strict '<' update, first arg-max. */
int p2c_fps(const float* xyz, const int64_t* start, int B, int N, int npoint,
            int64_t* out_idx /* (B,npoint) */, float* out_xyz /* (B,npoint,3) */, void* stream);

/* Ball query — replaces query_ball_point, models/pointnet_util.py:87-107 (and the (B,S,N)
 * square_distance :19-40 it thresholds).  First nsample indices, ascending, with
 * !(d > r2), d = ((-2*dot) + |q|^2) + |p|^2, dot = fma(q2,p2, fma(q1,p1, q0*p0)); padded with the
 * first hit; a query with no hit yields N everywhere (the reference's own out-of-range value).
 * r2 = (float)(radius*radius) computed by the caller in double then rounded (:102). */
int p2c_ball_query(const float* xyz /* (B,N,3) */, const float* new_xyz /* (B,S,3) */, int B, int N,
                   int S, float r2, int nsample, int64_t* out_idx /* (B,S,nsample) */, void* stream);

/* Grouping gather — replaces index_points + centring + concat in sample_and_group,
 * models/pointnet_util.py:130-139, and sample_and_group_all :146-163 (idx == NULL: every point in
 * order, one group per cloud, centre 0).  out row r = (b,s,j): [xyz[b,idx]-new_xyz[b,s] (3),
 * feats[b,idx] (D)], zero padded to ldo.  feats may be NULL (D = 0). */
int p2c_group(const float* xyz, const float* feats, int64_t ldf, const float* new_xyz,
              const int64_t* idx, int B, int N, int S, int nsample, int D, float* out, int64_t ldo,
              void* stream);

/* First MLP layer of a set-abstraction level fused with the grouping gather — replaces index_points + centring
 * + concat + the first Conv2d (models/pointnet_util.py:130-139, 200-201):
 *   Y[(b,s,j), c] = sum_d W[c,d] * (xyz[b,p,d] - new_xyz[b,s,d]) + Qf[b*N+p, c] + bias[c],  p = idx[b,s,j]
 * W: the conv weight (C, 3+D) with row stride ldw — only its first three (xyz) columns are read here; the
 * feature columns act through Qf = feats * W[:, 3:]^T (N rows per cloud, from p2c_linear; NULL when D = 0).
 * stats as in p2c_linear.  C is 64 or 128. */
int p2c_sa_first_layer(const float* xyz, const float* new_xyz, const int64_t* idx, const float* Qf, int64_t ldq,
                       const float* W, int64_t ldw, const float* bias, int B, int N, int S, int nsample, int C,
                       float* Y, int64_t ldy, double* stats, void* stream);

/* One 1x1-conv layer of a per-point MLP — replaces Conv2d/Conv1d (kernel 1) at
 * models/pointnet_util.py:200-203, :317-319 and models/pointnet_extrusion.py:58-65, with the
 * previous layer's BatchNorm+ReLU (and the head's dropout mask) folded into the operand load and
 * this layer's BatchNorm statistics and the nsample max-pool (:205) folded into the epilogue.
 *   A[m,k] = X[m,k]                                   (in_scale == NULL)
 *          = max(X[m,k]*in_scale[k]+in_shift[k], 0)   (otherwise)      [* in_mask[m,k] if given]
 *   Y[m,n] = sum_k A[m,k] * W[n,k] + bias[n]
 *   stats[n] += sum_m Y[m,n];  stats[N+n] += sum_m Y[m,n]^2          (stats != NULL, float64)
 *   pool_group G > 0: Ymax/Ymin[m/G, n] = max/min over the G consecutive rows of a group.
 * Y may be NULL when only the pooled output is wanted.  precision: P2C_PREC_*. */
#define P2C_PREC_FP32 0      /* SIMT fp32 FMA */
#define P2C_PREC_3XTF32 1    /* tcgen05 kind::tf32, error-compensated split (fp32-faithful) */
#define P2C_PREC_BF16 2      /* A BatchNorm whose finalisation is DEFERRED to the kernel that consumes the normalised values (p2c_linear,
 * p2c_pool_bn_relu, p2c_bn_relu_apply, p2c_head_masked take one as `in_bn` / `bn` in place of ready scale / shift
 * arrays): the consumer folds  scale = gamma / sqrt(var + eps),  shift = beta - mean * scale  per channel in its
 * prologue (same arithmetic as p2c_bn_finalize) and ONE of its CTAs writes scale_out / shift_out (required),
 * mean_out / invstd_out (optional, for the backward) and updates running_mean / running_var with `momentum`
 * (optional).  stats: float64 (2C) sum | sum of squares over `count` rows = batch statistics (train mode); NULL = use
 * running_mean / running_var (eval mode, nothing is updated).  Replaces nn.BatchNorm{1,2}d.forward,
 * models/pointnet_util.py:203, :318, without a launch of its own.  All pointers are device pointers. */
typedef struct p2c_bn_fold {
  const double* stats;
  int64_t count;
  const float* gamma;
  const float* beta;
  float eps;
  float momentum;
  float* running_mean;
  float* running_var;
  float* scale_out;
  float* shift_out;
  float* mean_out;
  float* invstd_out;
  int C;
} p2c_bn_fold;

/* tcgen05 kind::f16 with bf16 operands, fp32 accumulate */
int p2c_linear(const float* X, int64_t ldx, const float* W, const float* bias,
               const float* in_scale, const float* in_shift, const float* in_mask, int64_t ldmask,
               float* Y, int64_t ldy, int M, int N, int K, double* stats, int pool_group,
               float* Ymax, float* Ymin, int precision, const float* w_split, int64_t ldws,
               const p2c_bn_fold* in_bn /* NULL, or the pending BatchNorm of X (in_scale / in_shift then unused) */,
               void* stream);

/* A layer whose input is a concatenation [X | V broadcast over groups of rows] - PointNetFeaturePropagation with a
 * single source point per cloud (models/pointnet_util.py:298-299: `points2.repeat(1, N, 1)`, then the concat at :312
 * and the first Conv1d): by linearity  [X | V[g]] W^T + b = X W[:, :K]^T + (V W[:, K:]^T + b)[g],  g = row / group,
 * so the broadcast rows and the concat buffer never exist.
 * p2c_linear_small: Y = X W^T + bias for a handful of rows (M <= 256, K <= 3072; fp32 SIMT) - the per-group term.
 * p2c_linear_group_bias: the tcgen05 streamed-weight layer (3xTF32; w_split from p2c_split_tf32[_multi], which
 * takes a column slice of the weight through its source row stride) with bias_rows[(row / group), n] in place of a
 * per-channel bias; group % 32 == 0; in_scale / in_shift / in_bn and stats as in p2c_linear.
 * P2C_EUNSUPPORTED: shapes the streamed-weight kernel does not take (the caller then builds the concat). */
int p2c_linear_small(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* bias, float* Y,
                     int64_t ldy, int M, int N, int K, void* stream);
int p2c_linear_group_bias(const float* X, int64_t ldx, const float* w_split, int64_t ldws, const float* bias_rows,
                          int group, const float* in_scale, const float* in_shift, const p2c_bn_fold* in_bn /* or NULL */,
                          float* Y, int64_t ldy, int M, int N, int K, double* stats, void* stream);

/* The same first layer WITHOUT its output, for levels with no input features (D = 0: sa1 of the backbone,
 * models/pointnet_util.py:130-139, 200-203 and the layer after it).  The raw first-layer rows (B*S*nsample, C0) are
 * neither written nor read back: p2c_sa_xyz_linear is the level's SECOND layer (p2c_linear semantics for W1 / b1 / Y /
 * stats / pool_group / Ymax / Ymin, tensor cores, 3xTF32) whose operand transform recomputes row (b,s,j) of its input
 * as max(scale0 * (W0 (xyz[b,idx] - new_xyz[b,s]) + b0) + shift0, 0) - with scale0 / shift0 folded into the conv's
 * coefficients, three FMAs per element (within one rounding of the materialised p2c_sa_first_layer + BatchNorm).  Train-mode BatchNorm statistics of the first layer come in closed form from nine
 * moments of the centred neighbour coordinates: p2c_group_moments (coordinates only; `moments` holds
 * p2c_group_moments_size() doubles: per-CTA partials, the nine sums, and a launch counter in the last word that must
 * be ZERO before the first launch and is left zero by every launch).  Given `moments` and a pending train-mode bn0,
 * p2c_sa_xyz_linear derives the first layer's sum / sum-of-squares itself (and stores them into bn0->stats);
 * p2c_sa_xyz_stats is the same closed form as a stand-alone kernel (writes the 2C sums where p2c_sa_first_layer
 * would have accumulated them).  scale0 / shift0 or bn0 as in p2c_linear.
 * Returns P2C_EUNSUPPORTED for shapes the tcgen05 kernel does not take (the caller then runs p2c_sa_first_layer).
 * A 64 -> 64 second layer whose rows are kept (C0 = N1 = 64, pool_group = 0, Y != NULL: sa1.1 of the backbone) runs on the
 * rows-as-lanes kernel of csrc/sa_stack_tc.cu - conv0, conv1 and its bias on the tensor core, every A operand in tensor
 * memory - same semantics, results within fp32 rounding of the other kernel (env P2C_SA_PAIR=0 selects that one). */
int p2c_group_moments_size(void);
int p2c_group_moments(const float* xyz, const float* new_xyz, const int64_t* idx, int B, int N, int S, int nsample,
                      double* moments, void* stream);
int p2c_sa_xyz_stats(const double* moments, int64_t rows, const float* W, int64_t ldw, const float* bias, int C,
                     double* stats, void* stream);
int p2c_sa_xyz_linear(const float* xyz, const float* new_xyz, const int64_t* idx, int B, int N, int S, int nsample,
                      const float* W0, int64_t ldw0, const float* b0, int C0, const float* scale0, const float* shift0,
                      const p2c_bn_fold* bn0 /* or NULL */, const double* moments /* or NULL */, const float* W1,
                      const float* b1, int N1, float* Y,
                      int64_t ldy, double* stats, int pool_group, float* Ymax, float* Ymin, void* stream);

/* A WHOLE feature-less set-abstraction level as ONE kernel (csrc/sa_stack_tc.cu; models/pointnet_util.py:130-139,
 * 181-207 for sa1: gather, Conv2d 3 -> C0 + BN + ReLU, C0 -> C1 + BN + ReLU, C1 -> C2 + BN + ReLU, max over nsample),
 * for BatchNorm layers whose statistics are known before the launch: eval mode (running statistics, bn->stats NULL;
 * eval.py) or caller-supplied sums.  Reads only xyz / new_xyz / idx and the weights, writes only the pooled
 * post-BN/ReLU rows out (B*S, C2): no activation of the level crosses HBM.  The two 64-wide contractions run on tcgen05
 * (3xTF32: fp32-faithful), chained through shared memory (rows-as-lanes for C0 -> C1 with both operands in shared
 * memory, channels-as-lanes with the weights in tensor memory for C1 -> C2, whose epilogue pools in registers).
 * bn0 / bn1 / bn2: the three BatchNorm descriptors (scale_out / shift_out are written as by every folding consumer).
 * C0 = C1 = 64, C2 <= 128, nsample in {32, 64, 128}; P2C_EUNSUPPORTED otherwise (the caller runs the per-layer path). */
int p2c_sa_stack_fused(const float* xyz, const float* new_xyz, const int64_t* idx, int B, int N, int S, int nsample,
                       const float* W0, int64_t ldw0, const float* b0, const p2c_bn_fold* bn0, const float* W1,
                       const float* b1, const p2c_bn_fold* bn1, const float* W2, const float* b2,
                       const p2c_bn_fold* bn2, int C0, int C1, int C2, float* out, int64_t ldo, void* stream);


/* hi/lo tf32 split of a weight matrix for the large-K tensor-core kernel: out[0][n][k] = w with the low 13
 * mantissa bits cleared, out[1][n][k] = w - hi; rows padded with zeros to ldw (multiple of 4) floats.
 * Pass the result as w_split/ldws to p2c_linear; NULL keeps large-K layers on the fp32 SIMT kernel. */
int p2c_split_tf32(const float* W, int N, int K, float* out /* (2,N,ldw) */, int64_t ldw, void* stream);
/* The same for up to 16 weight matrices in ONE launch (host arrays of `count` device pointers / sizes): every
 * streamed-weight layer of a forward pass is split by a single kernel.  transposed (NULL = all 0): entry i != 0 means
 * W[i] is stored as (K, N) and the split of its TRANSPOSE (N, K) is written (the reverse sweeps multiply by W^T).
 * src_ld (NULL = dense): row stride of W[i] as stored, so a column block of a wider matrix can be the source. */
int p2c_split_tf32_multi(const float* const* W, const int* N, const int* K, float* const* out, const int64_t* ldw,
                         const int* transposed, const int64_t* src_ld, int count, void* stream);

/* One hidden layer of the implicit sketch network (IGR/network.py:8-92: ImplicitNet.forward and gradient()) on the
 * tensor cores, 3xTF32, weights pre-split (p2c_split_tf32[_multi]):
 *   op 1 (forward sweep):  Z = X W^T + bias;  Y = softplus_beta(Z) * oscale  (nn.Softplus(beta), threshold 20);
 *                          S = sigmoid(beta Z) (1 above the threshold) when S != NULL - softplus'(Z), kept for the
 *                          closed-form input gradient;
 *   op 2 (reverse sweep):  Y = (X W^T + bias) * Mul * oscale, Mul (M, N) read element-wise: a_{i-1} = s_{i-1} * (a_i W_i);
 *   op 0:                  Y = X W^T + bias.
 * X (M, K) rows with stride ldx (16-byte aligned rows), Y / S / Mul (M, N) rows; channels >= N are not written. */
int p2c_linear_act(const float* X, int64_t ldx, const float* w_split, int64_t ldws, const float* bias, int M, int N,
                   int K, int op, float beta, float oscale, float* Y, int64_t ldy, float* S, int64_t lds,
                   const float* Mul, int64_t ldmul, void* stream);

/* One layer-GEMM of the implicit network's SECOND-ORDER backward (training through gradient(..., create_graph=True),
 * IGR/network.py:8-17 + train_Point2Cyl.py:608-672, restated without autograd in oracle/igr_oracle.py
 * implicit_backward_closed_form).  Same tensor-core kernel as p2c_linear_act, no bias, two more element-wise epilogues:
 *   op 3 (adjoint of the reverse sweep, going up):  acc = X W^T with X = r_bar_i, W = W_i;  Y = acc * Mul * oscale is the
 *         next layer's r_bar (Mul = softplus'(z_i));  Z = beta * acc * V * (1 - Mul) with V = a_i is the extra
 *         pre-activation gradient a_bar_i * q_{i+1} * softplus''(z_i)  (softplus'' = beta s (1 - s), a_i = s_i q_{i+1});
 *   op 4 (backward of the forward sweep, going down):  Y = acc * Mul * oscale + V with X = delta_i, W = W_i^T,
 *         Mul = softplus'(z_{i-1}), V = that extra term of layer i-1 (Y may alias V).
 * Mul / V / Y / Z are (M, N) row matrices; Z is written for op 3 only. */
int p2c_linear_act_bwd(const float* X, int64_t ldx, const float* w_split, int64_t ldws, int M, int N, int K, int op,
                       float beta, float oscale, float* Y, int64_t ldy, const float* Mul, int64_t ldmul, const float* V,
                       int64_t ldv, float* Z, int64_t ldz, void* stream);

/* bf16 copy of a weight matrix for P2C_PREC_BF16: out (N, ldw) bf16 (round-to-nearest), rows zero padded to ldw
 * (multiple of 8) elements.  Pass it as w_split (ldws = ldw) to p2c_linear with precision P2C_PREC_BF16: the layer then
 * runs ONE tcgen05 kind::f16 pass on bf16 operands with fp32 accumulation (not fp32-faithful; BASELINE.json's bf16
 * MLP-stack configuration). */
int p2c_cast_bf16(const float* W, int N, int K, void* out /* (N,ldw) bf16 */, int64_t ldw, void* stream);

/* Which kernel p2c_linear dispatches a (16-byte aligned) layer to: 0 = fp32 SIMT, 1 = tcgen05 3xTF32 with the
 * weights resident in tensor memory (K <= 192), 2 = tcgen05 3xTF32 with streamed weights (needs w_split),
 * 3 = tcgen05 bf16 with streamed weights (P2C_PREC_BF16, needs the p2c_cast_bf16 copy as w_split).
 * Pure function of the shape; lets tests assert that the tensor-core path really ran. */
int p2c_linear_path(int64_t ldx, int M, int N, int K, int has_mask, int pool_group, int precision, int has_split);

/* Tools only (tools/tc_timeline.py): device buffer of 4*512 int64 that CTA 0 of subsequent tensor-core
 * launches fills with (tag, globaltimer) pairs per warp role; NULL (default) disables the probes. */
int p2c_debug_set_timeline(void* buf);

/* Output heads — replaces F.dropout + the fc2 Conv1d heads, models/pointnet_extrusion.py:60-65:
 *   Y[b*N+n, j] = sum_k max(H[b*N+n,k]*scale[k]+shift[k], 0) * mask_cf[b,k,n] * W[j,k] + bias[j]
 * H: fc1's raw output (B*N, C) rows; scale/shift: folded bn1 (NULL = identity, no ReLU); mask_cf: the
 * (B, C, N) channel-first multiplicative dropout mask exactly as F.dropout(ones(B,C,N)) returns it, or NULL;
 * dropout_seed (used when mask_cf is NULL): device pointer to two int64 words; the kernel then draws the p = 0.5
 * mask itself - keep-bit of (point m, channel c) = one bit of Philox4x32-10(counter {m, c/128, seed[1]}, key seed[0]),
 * kept values scaled by 2 - so no (B, C, N) mask tensor crosses HBM (F.dropout semantics, its own random stream);
 * p2c_head_bwd regenerates the same bits from the same two words.  Both NULL = no dropout.
 * W: the heads' weights concatenated (Nout, C), Nout <= 36, C <= 256, C % 16 == 0.
 * precision: P2C_PREC_3XTF32 with a folded BatchNorm and no explicit mask runs on the tcgen05 layer kernel (the mask
 * is drawn in its operand transform: same bits); otherwise (explicit mask, P2C_PREC_FP32, shapes the tensor-core
 * kernel does not take) the fp32 SIMT kernel. */
int p2c_head_masked(const float* H, int64_t ldh, const float* scale, const float* shift,
                    const float* mask_cf, const int64_t* dropout_seed, const float* W, const float* bias, float* Y,
                    int64_t ldy, int B, int N, int C, int Nout, const p2c_bn_fold* bn /* or NULL */, int precision,
                    void* stream);

/* BatchNorm bookkeeping — replaces the statistics half of nn.BatchNorm{1,2}d (eps, momentum,
 * unbiased running_var) used at models/pointnet_util.py:201-203, :317-319, pointnet_extrusion.py:59.
 * training != 0: mean/var from stats (count rows), running stats updated in place with `momentum`;
 * training == 0: running stats.  Writes scale = gamma/sqrt(var+eps), shift = beta - mean*scale. */
int p2c_bn_finalize(const double* stats, int64_t count, const float* gamma, const float* beta,
                    float eps, float momentum, int training, float* running_mean,
                    float* running_var, float* scale, float* shift, float* save_mean,
                    float* save_invstd, int C, void* stream);

/* out[m,c] = max(Y[m,c]*scale[c]+shift[c], 0) — the BN+ReLU application where a layer's output
 * has to exist in memory (module outputs). */
int p2c_bn_relu_apply(const float* Y, int64_t ldy, const float* scale, const float* shift,
                      float* out, int64_t ldo, int64_t M, int C, const p2c_bn_fold* bn /* or NULL */, void* stream);

/* Pooled BN+ReLU: out[g,c] = max(v*scale[c]+shift[c], 0) with v = scale[c] >= 0 ? Ymax : Ymin —
 * equals max over the group of relu(bn(y)) (models/pointnet_util.py:203-205) by monotonicity.
 * Also emits nothing else; arg-max routing for backward is recomputed there. */
int p2c_pool_bn_relu(const float* Ymax, const float* Ymin, const float* scale, const float* shift,
                     float* out, int64_t ldo, int64_t G, int C, const p2c_bn_fold* bn /* or NULL */, void* stream);

/* 3-NN inverse-distance interpolation — replaces PointNetFeaturePropagation's
 * square_distance + sort + gather + weighted sum, models/pointnet_util.py:301-308.
 * Distances in the reference's expanded form (may be slightly negative), w = 1/(d+1e-8) normalised.
 * out row (b,n) gets D floats at out + row*ldo.  idx_out (B,N,3) int64 / w_out (B,N,3) optional. */
int p2c_three_nn_interp(const float* xyz1 /* (B,N,3) */, const float* xyz2 /* (B,S,3) */,
                        const float* feats2 /* (B*S, D) */, int64_t ldf, int B, int N, int S, int D,
                        float* out, int64_t ldo, int64_t* idx_out, float* w_out, void* stream);

/* The two halves of p2c_three_nn_interp on their own.  The neighbour search depends on the coordinates only, the
 * gather on the features only, so a pipelined caller runs them in different stages:
 *   p2c_three_nn_search: idx (B,N,3) int64 and w (B,N,3) of the three nearest sources (same arithmetic, same ties);
 *   p2c_three_nn_gather: out[(b,n), :] = sum_j w[b,n,j] * feats2[b*S + idx[b,n,j], :]  (same rounding order). */
int p2c_three_nn_search(const float* xyz1, const float* xyz2, int B, int N, int S, int64_t* idx_out, float* w_out,
                        void* stream);
int p2c_three_nn_gather(const float* feats2, int64_t ldf, const int64_t* idx, const float* w, int B, int N, int S, int D,
                        float* out, int64_t ldo, void* stream);

/* Loss pass 1 — one sweep over the points that produces every per-cloud sufficient statistic of
 * train_Point2Cyl_without_sketch.py:246-353: unit normals (:247), softmax over 2K and the
 * barrel/base split (:254-265), normal loss sum (losses.py:130), the Hungarian cost ingredients
 * (losses.py:38-41), centre sums (data_utils.py:253-266) and the 3x3 scatter matrices of
 * estimate_extrusion_axis (data_utils.py:155-163) for every predicted column.
 * stats: (B, P2C_SEG_STRIDE(K)) float32, layout in point2cyl_b200/csrc/segfit.cu. */
int p2c_segfit_stats_stride(int K);
int p2c_segfit_stats(const float* X_raw, int64_t ldx, const float* W_raw, int64_t ldw,
                     const float* pcs, const float* gt_normals, const int64_t* inst,
                     const int64_t* bb, int B, int N, int K, float* partial /* scratch */,
                     int64_t partial_elems /* >= B*ceil(N/256)*stride */, float* stats, void* stream);

/* Same statistics from soft assignments the caller already holds (function-level losses.py / data_utils.py
 * API: hungarian_matching, compute_miou_loss, estimate_extrusion_axis, estimate_extrusion_centers).
 * wb/wc: (B,N,K), row stride ld*, element stride s* (slices such as W_2K[:, :, ::2] are operands); wc, X, pcs,
 * gt_normals, inst, bb may each be NULL (their statistics are then zero / labels -1). */
int p2c_segfit_stats_w(const float* X, int64_t ldx, int normalize_x, const float* wb, int64_t ldb, int64_t sb,
                       const float* wc, int64_t ldc, int64_t sc, const float* pcs, const float* gt_normals,
                       const int64_t* inst, const int64_t* bb, int B, int N, int K, float* partial,
                       int64_t partial_elems, float* stats, void* stream);

/* Hungarian cost (losses.py:39-42): cost (B,K,K) = D / max(cnt_g + colsum_k - D, 1e-10), n_gt (B). */
int p2c_segfit_cost(const float* stats, int B, int K, float* cost, int32_t* n_gt, void* stream);

/* On-device assignment — replaces scipy.optimize.linear_sum_assignment(-cost) at losses.py:43-45:
 * match[b, g] = column assigned to gt row g < n_gt[b] (maximising the summed score), 0 for g >= n_gt[b].
 * Exact optimum (Hungarian algorithm, float64); K <= 16. */
int p2c_hungarian(const float* score /* (B,K,K) */, const int32_t* n_gt /* (B) */, int B, int K,
                  int64_t* match /* (B,K) */, void* stream);

/* Loss pass 2 — the base/barrel loss of train_Point2Cyl_without_sketch.py:283-313 given the match,
 * summed per cloud (the sort at :292 cancels out of the sum). bb_sum: (B). */
int p2c_bb_loss(const float* W_raw, int64_t ldw, const int64_t* bb, const int64_t* match,
                const int32_t* n_gt, int B, int N, int K, float* partial /* scratch */,
                int64_t partial_elems /* >= B*ceil(N/256) */, float* bb_sum, void* stream);

/* Loss finalisation — per (cloud, gt slot): relaxed IoU (losses.py:95-101), centre
 * (data_utils.py:253-266), axis = eigenvector of the smallest eigenvalue of the matched 3x3
 * scatter (data_utils.py:162-172; cyclic Jacobi, one warp per segment), then the masked means of
 * losses.py:83-88 / train_...:330-352.  out_losses: 6 floats {total, normal, miou, bb, axis, center}. */
int p2c_loss_finalize(const float* stats, const float* bb_sum, const int64_t* match,
                      const int32_t* n_gt, const float* gt_axes, const float* gt_centers, int B,
                      int N, int K, int norm_eig, const float* weights /* host, 5 */,
                      float* E_AX /* (B,K,3) */, float* centers /* (B,K,3) */,
                      float* per_seg /* (B,K,3) {1-iou, axis, centre} */,
                      float* per_cloud /* (B,5) {miou, normal, bb, axis, centre} */,
                      float* out_losses /* (6) */, void* stream);

/* Smallest-eigenvalue eigenvector of n symmetric 3x3 matrices (row-major 9 floats each) —
 * replaces torch.symeig(...)[1][:, :, 0], data_utils.py:170-171. */
int p2c_eig3x3_smallest(const float* M, int n, float* vec /* (n,3) */, float* eval /* (n,3) */,
                        void* stream);

/* Function-level helpers of the reference's module API, kept for completeness:
 * square_distance (models/pointnet_util.py:19-40; same rounding as the ball query) and
 * index_points (:43-60; out[b,m,:] = points[b, idx[b,m], :], points rows have stride ldp). */
int p2c_square_distance(const float* src /* (B,S,3) */, const float* dst /* (B,N,3) */, int B, int S,
                        int N, float* out /* (B,S,N) */, void* stream);
int p2c_gather_rows(const float* points, int64_t ldp, const int64_t* idx /* (B,Mper) */, int B, int N,
                    int Mper, int C, float* out /* (B,Mper,C) */, void* stream);

/* ---- second wave: projection / scale / extent closed forms and eval helpers (a18) ---- */

/* Order-preserving member lists per (cloud, segment) — replaces the `one_hot` / `where(bb==0)` / `nonzero()`
 * selection of data_utils.py:1018-1061 (and :1654-1695): point n is a member of list (b,k) iff
 * seg_label[b,n] == k and (bb == NULL or bb[b,n] == bb_value).  counts (B,K) int32; lists (B,K,N) int32, first
 * counts[b,k] entries valid, ascending point index (the order nonzero() yields); lists may be NULL (counts only). */
int p2c_segment_lists(const int64_t* seg_label, const int64_t* bb, int bb_value, int B, int N, int K,
                      int32_t* counts, int32_t* lists, void* stream);

/* Sketch projection — replaces the body of sketch_implicit_projection / 2 / 3, data_utils.py:1014-1417:
 * per (segment k, cloud b) gather the sampled member points (member number rand_idx[k,b,s] of list (b,k);
 * rand_idx == NULL: member s itself, lists == NULL: every point is a member — variant 3), rotate the extrusion axis
 * onto +z with torchgeometry's angle_axis_to_rotation_matrix applied to (a x z)*angle (the axis is NOT normalised
 * upstream, :1096-1103 — replicated), drop z, subtract the projected centre (:1127-1131), scale = max 2-norm
 * (:1136).  Segments with <= 1 member over the batch give zeros / scale 1 (:1042-1044); clouds with <= 1 member
 * give -centre_projected / scale 1 (:1054-1056 leaves zero rows that are still centred).  X / X_proj may be NULL.
 * P_proj, X_proj: (K,B,S,2); scales (K,B); found (B,K) 0/1 or NULL. */
int p2c_sketch_project(const float* P, const float* X, int B, int N, int K, int S, const int32_t* lists,
                       const int32_t* counts, const int64_t* rand_idx /* (K,B,S) */, const float* axes /* (B,K,3) */,
                       const float* centers /* (B,K,3) */, float zero_tol, float* P_proj, float* X_proj,
                       float* scales, float* found, int32_t* sel_out /* (K,B,S) sampled point or -1, may be NULL */,
                       float* R_out /* (K,B,9) rotation, may be NULL */, void* stream);

/* Backward of the projected NORMALS w.r.t. X (the with-sketch trainer feeds predicted normals,
 * train_Point2Cyl.py:549): dX[b, sel[k,b,s], :] += R[k,b][:, 0:2] * dX_proj[k,b,s,:].  dX (B,N,3) is overwritten. */
int p2c_sketch_project_bwd(const float* dX_proj, const int32_t* sel, const float* R, int B, int N, int K, int S,
                           float* dX, void* stream);

/* Extents along the axis — replaces get_extrusion_extents, data_utils.py:1650-1730: min / max over the sampled
 * members of (p - c).a; same selection and not-found conventions as above.  extents (K,B,2). */
int p2c_extrusion_extents(const float* P, int B, int N, int K, int S, const int32_t* lists, const int32_t* counts,
                          const int64_t* rand_idx, const float* axes, const float* centers, float* extents,
                          float* found, void* stream);

/* hard_W_encoding, losses.py:55-68: hard[b,n,:] = one-hot(first arg-max_k W[b,n,k]), zeroed when that column's
 * sum over N (colsum (B,K), from p2c_segfit_stats_w; NULL = keep all) is below null_below = N * threshold.
 * W rows have stride ldw and element stride sw.  label (B,N) int64 = the arg-max (eval.py:328, :339); either output
 * may be NULL. */
int p2c_hard_w_encoding(const float* W, int64_t ldw, int64_t sw, int B, int N, int K, const float* colsum,
                        float null_below, float* hard /* (B,N,K) */, int64_t* label /* (B,N) */, void* stream);

/* compute_normal_difference / acos_safe, losses.py:123-124, :146-159: scale * acos(clamp(|<x,g>|, +-(1-1e-6))) per
 * point (per_point (B,N) or NULL) and summed per cloud (per_cloud_sum (B) or NULL). */
int p2c_normal_angle(const float* X, const float* G, int B, int N, float scale, float* per_point,
                     float* per_cloud_sum, void* stream);

/* ---- backward of the forward+loss path (the training step) ---- */

/* d total / d statistics for the loss block — backward of p2c_loss_finalize (relaxed IoU losses.py:95-101,
 * centre data_utils.py:253-266, axis data_utils.py:162-172 through the eigenvector sensitivity that
 * torch.symeig's backward implements, normal loss losses.py:130).  eff_weights: DEVICE pointer to 5 floats
 * {seg, normal, bb, axis, centre} = loss multipliers times the upstream gradient.  dstats: (B, stride(K)),
 * fully overwritten. */
int p2c_loss_backward_coef(const float* stats, const int64_t* match, const int32_t* n_gt, const float* gt_axes,
                           const float* gt_centers, const float* eff_weights, int B, int N, int K, int norm_eig,
                           float* dstats, void* stream);

/* Backward of p2c_segfit_stats (+ p2c_bb_loss when match/n_gt are given): one sweep over the points from the
 * statistics' gradient to the network outputs — through W = barrel + base, the squared weights of the 3x3
 * scatters, softmax over 2K (train_...:254), F.normalize (:247) and the base/barrel cross entropy (:286-307).
 * dX_raw (B*N rows, stride lddx >= 3), dW_raw (B*N rows, stride lddw >= 2K): both fully written. */
int p2c_segfit_backward(const float* X_raw, int64_t ldx, const float* W_raw, int64_t ldw, const float* pcs,
                        const float* gt_normals, const int64_t* inst, const int64_t* bb, const float* dstats,
                        const int64_t* match, const int32_t* n_gt, const float* eff_weights, int B, int N, int K,
                        float* dX_raw, int64_t lddx, float* dW_raw, int64_t lddw, void* stream);

/* Backward of p2c_segfit_stats_w (function-level losses.py / data_utils.py API): dX (rows, stride lddx) or NULL,
 * dWb / dWc contiguous (B,N,K) or NULL. */
int p2c_segfit_backward_w(const float* X, int64_t ldx, int normalize_x, const float* wb, int64_t ldb, int64_t sb,
                          const float* wc, int64_t ldc, int64_t sc, const float* pcs, const float* gt_normals,
                          const int64_t* inst, const float* dstats, int B, int N, int K, float* dX, int64_t lddx,
                          float* dWb, float* dWc, void* stream);

/* Backward of p2c_eig3x3_smallest: dM (n,3,3) symmetric from gvec = d L / d vec (n,3) — what autograd does for
 * torch.symeig(...)[1][:, :, 0] at data_utils.py:170-171. */
int p2c_eig3x3_backward(const float* M, const float* gvec, int n, float* dM, void* stream);

/* BatchNorm+ReLU backward, pass 1 — autograd of models/pointnet_util.py:203 / :319 (F.relu(bn(conv(x)))):
 * sums[c] = sum_m g, sums[C+c] = sum_m g*Y[m,c] with g = dA[m,c] * [scale[c]*Y[m,c]+shift[c] > 0]
 * (scale == NULL: no ReLU gate).  sums: float64 (2C), zeroed here.  C: 32..1024, power of two. */
int p2c_bn_bwd_reduce(const float* dA, int64_t ldda, const float* Y, int64_t ldy, const float* scale,
                      const float* shift, int64_t M, int C, double* sums, void* stream);

/* Same sums for a max-pooled layer (torch.max(new_points, 2)[0], models/pointnet_util.py:205): only the arg-max
 * row of each group carries gradient and its raw value is Ymax (Ymin when scale < 0), so the pooled tensors
 * suffice.  dOut: (G, C) gradient of the pooled post-BN/ReLU output. */
int p2c_pool_bwd_reduce(const float* dOut, int64_t ldd, const float* Ymax, const float* Ymin, const float* scale,
                        const float* shift, int64_t G, int C, double* sums, void* stream);

/* BatchNorm backward, per-channel part: dgamma[c] += invstd*(s2 - mean*s1), dbeta[c] += s1 (either may be NULL) and
 * coef (3C) = {a, b, c} with dY = a*g + b*Y + c (training: batch statistics over `count` rows; eval: b = c = 0). */
int p2c_bn_bwd_coef(const double* sums, int64_t count, const float* gamma, const float* mean, const float* invstd,
                    int training, float* coef, float* dgamma, float* dbeta, int C, void* stream);

/* BatchNorm+ReLU backward, pass 2: dY[m,c] = a*g + b*Y + c (may run in place over dA). */
int p2c_bn_bwd_apply(const float* dA, int64_t ldda, const float* Y, int64_t ldy, const float* scale,
                     const float* shift, const float* coef, int64_t M, int C, float* dY, int64_t lddy, void* stream);

/* Pass 2 for a max-pooled layer: dY (G*group, C) from dOut (G, C); the gradient is routed to the first row of the
 * group whose raw value equals the pooled one. */
int p2c_pool_bwd_apply(const float* dOut, int64_t ldd, const float* Ymax, const float* Ymin, const float* Y,
                       int64_t ldy, const float* scale, const float* shift, const float* coef, int64_t G, int group,
                       int C, float* dY, int64_t lddy, void* stream);

/* p2c_bn_bwd_coef folded into the kernel that applies it (one launch per layer instead of two; same float64
 * coefficient arithmetic, evaluated per thread for its own channels): dgamma / dbeta are accumulated by exactly one
 * thread per channel.  p2c_bn_bwd_apply_fused needs C / 4 to divide 256 or be a multiple of it (P2C_EUNSUPPORTED
 * otherwise: run the two-kernel form). */
int p2c_bn_bwd_apply_fused(const float* dA, int64_t ldda, const float* Y, int64_t ldy, const float* scale,
                           const float* shift, const double* sums, int64_t count, const float* gamma,
                           const float* mean, const float* invstd, int training, float* dgamma, float* dbeta,
                           int64_t M, int C, float* dY, int64_t lddy, void* stream);
int p2c_pool_bwd_apply_fused(const float* dOut, int64_t ldd, const float* Ymax, const float* Ymin, const float* Y,
                             int64_t ldy, const float* scale, const float* shift, const double* sums, int64_t count,
                             const float* gamma, const float* mean, const float* invstd, int training, float* dgamma,
                             float* dbeta, int64_t G, int group, int C, float* dY, int64_t lddy, void* stream);

/* Weight / bias gradient of a 1x1 conv — autograd of Conv2d/Conv1d at models/pointnet_util.py:201, :317,
 * pointnet_extrusion.py:58-65:  dW[n,k] += sum_m dY[m,n] * A[m,k], db[n] += sum_m dY[m,n], with the layer input
 * A recomputed from the raw previous activation exactly as p2c_linear's operand load does
 * (A = max(X*in_scale+in_shift, 0) [* mask_cf[b,k,n], m = b*mask_N + n]).  dW/db are accumulated (atomics).
 * P2C_PREC_3XTF32: split-over-rows tcgen05 kernel (csrc/wgrad_tc.cu: MN-major operands straight from TMA boxes,
 * accumulator in tensor memory, error-compensated tf32 like the forward); masked / unaligned / tiny calls and
 * P2C_PREC_FP32 take the register-tiled SIMT kernel. */
int p2c_wgrad(const float* dY, int64_t lddy, const float* X, int64_t ldx, const float* in_scale,
              const float* in_shift, const float* mask_cf, int mask_N, int64_t M, int N, int K, float* dW,
              int64_t lddw, float* db, int precision /* P2C_PREC_FP32: SIMT; P2C_PREC_3XTF32: tcgen05 */,
              void* stream);

/* 1 when p2c_wgrad(P2C_PREC_3XTF32) runs on the tensor cores for these strides / sizes (16-byte aligned operands
 * assumed), 0 when it takes the fp32 SIMT kernel.  Lets tests assert that the tcgen05 path really ran. */
int p2c_wgrad_path(int64_t lddy, int64_t ldx, int64_t M, int N, int K, int has_mask);

/* Backward of p2c_sa_first_layer: dQf[b*N+p, :] += dY[r, :] (NULL when the level has no input features; pre-zeroed
 * by the caller), dW[:, 0:3] += dY^T * (xyz[p] - centre), dbias += column sums. */
int p2c_sa_first_bwd(const float* dY, int64_t lddy, const float* xyz, const float* new_xyz, const int64_t* idx, int B,
                     int N, int S, int nsample, int C, float* dQf, int64_t ldq, float* dW, int64_t lddw, float* dbias,
                     void* stream);

/* Backward of p2c_group (the un-fused grouping path): dfeats[b*N + idx[b,s,j], :] += dRows[(b,s,j), 3:3+D]; dfeats must
 * be zeroed by the caller; the three xyz columns carry no gradient (coordinates are data). */
int p2c_group_bwd(const float* dRows, int64_t ldr, const int64_t* idx, int B, int N, int S, int nsample, int D,
                  float* dfeats, int64_t ldf, void* stream);

/* Backward of p2c_three_nn_interp: dfeats2[b*S + idx[b,n,j], :] += w[b,n,j] * dInterp[b*N+n, :] (S == 1: the
 * broadcast of models/pointnet_util.py:298-299, idx/w unused).  dfeats2 is fully defined on return. */
int p2c_three_nn_interp_bwd(const float* dInterp, int64_t ldi, const int64_t* idx, const float* w, int B, int N, int S,
                            int D, float* dfeats2, int64_t ldf, void* stream);

/* Data gradient of p2c_head_masked: dA[m,k] = mask_cf[b,k,n] * sum_j dOut[m,j] * W[j,k] (the ReLU/BN part is then the
 * generic p2c_bn_bwd_* on fc1's raw output).  With A_out != NULL the kernel also writes the heads' input
 * A[m,k] = max(H[m,k]*scale[k]+shift[k], 0) * mask_cf[b,k,n] (H: fc1's raw output), so that the heads' dW/db come
 * from the tensor-core p2c_wgrad on a plain row matrix. */
int p2c_head_bwd(const float* dOut, int64_t ldo, const float* mask_cf, const int64_t* dropout_seed, const float* W,
                 int B, int N, int C, int Nout, float* dA, int64_t ldda, const float* H, int64_t ldh, const float* scale,
                 const float* shift, float* A_out, int64_t lda, void* stream);

/* torch.optim.Adam step (train_Point2Cyl_without_sketch.py:189, :368) over a flat fp32 parameter buffer:
 * g = grads*grad_scale (+ weight_decay*p), m/v updated in place, bias-corrected with `step` (1-based). */
int p2c_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);

/* ---- implicit sketch network of the with-sketch trainer -------------------------------------------
 * add_latent (IGR/network.py:200-206): X0[r] = [latent[r / S] (E) | pts[r] (2)], zero padded to ldx columns; when P is
 * given the same row, times pscale, is also written at column colp of P (the skip concatenation, network.py:80-81). */
int p2c_igr_add_latent(const float* latent /* (R/S, E) */, const float* pts /* (R, 2) */, int64_t R, int S, int E,
                       float* X0, int64_t ldx, float* P, int64_t ldp, int colp, float pscale, void* stream);
/* out[m, c] = A[m, c] * w[c]: seed of the closed-form input-gradient sweep, a = softplus'(z) * W_last. */
int p2c_igr_scale_cols(const float* A, int64_t lda, const float* w, int64_t M, int C, float* out, int64_t ldo,
                       void* stream);
/* out[m, j] (+)= scale * <A[m, :C], V[j, :C]> + bias[j], NV in {1, 2}: the 512 -> 1 output layer (network.py:88-91) and
 * the two columns of d f / d x that gradient() keeps (network.py:17). */
int p2c_igr_rowdots(const float* A, int64_t lda, int64_t M, int C, const float* V, int64_t v_stride_j,
                    int64_t v_stride_c /* V[j, c] = V[j * v_stride_j + c * v_stride_c] */, int NV, const float* bias,
                    float scale, float* out, int64_t ldo, int accumulate, void* stream);
/* Per sketch instance i (S on-surface rows, S_off off-surface rows): out[i] = { mean |f_on|, mean min(|g_on - n|,
 * |g_on + n|), mean (|g_off| - 1)^2 } - the manifold, SALD-normal and eikonal terms before the masked instance mean,
 * train_Point2Cyl.py:627-647. */
int p2c_igr_loss_terms(const float* f_on, const float* g_on, const float* normals, const float* g_off, int instances,
                       int S, int S_off, float* out /* (instances, 3) */, void* stream);
/* ---- its backward (oracle/igr_oracle.py implicit_backward_closed_form; the layer GEMMs are p2c_linear_act_bwd and
 * p2c_wgrad) ----
 * out[j, c] += scale * sum_m G[m, j] * A[m, c], NV in {1, 2} rows of out; G NULL (NV = 1) = plain column sums: the
 * 512 -> 1 output layer's weight gradient from both sweeps. */
int p2c_igr_colsums(const float* A, int64_t lda, int64_t M, int C, const float* G, int64_t ldg, int NV, float scale,
                    float* out, int64_t ldo, void* stream);
/* out[m, c] = ZE[m, c] + S[m, c] * f_bar[m] * w[c]: pre-activation gradient of the last hidden layer (the output
 * layer's data gradient is an outer product).  ZE NULL = no input-gradient loss, f_bar NULL = no value loss. */
int p2c_igr_seed_delta(const float* ZE, int64_t ldz, const float* S, int64_t lds, const float* f_bar, const float* w,
                       int64_t M, int C, float* out, int64_t ldo, void* stream);
/* Backward of add_latent over one block of instances x S rows: dlatent[i, :E] += scale * sum_p DX[i*S+p, :E];
 * dpts[r, 0..1] (+)= scale * DX[r, E..E+1] when dpts != NULL. */
int p2c_igr_latent_grad(const float* DX, int64_t lddx, int instances, int S, int E, float scale, float* dlatent,
                        float* dpts, int accumulate_pts, void* stream);
/* Backward of p2c_igr_loss_terms: dterms (instances, 3) -> f_bar (instances*S), g_bar_on (instances*S, 2),
 * g_bar_off (instances*S_off, 2), with torch's sub-gradient conventions (sign(0) = 0, d||v||/dv = 0 at v = 0, torch.min
 * routes a tie to its first argument |g - n|). */
int p2c_igr_loss_terms_bwd(const float* f_on, const float* g_on, const float* normals, const float* g_off, int instances,
                           int S, int S_off, const float* dterms, float* f_bar, float* g_bar_on, float* g_bar_off,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* POINT2CYL_H */
