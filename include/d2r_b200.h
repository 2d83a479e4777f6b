/* d2r_b200 -- C ABI of the B200-native routed interaction stack.
 *
 * Drop-in boundary for D2R's dual-branch dynamic-routing interaction stack
 * (reference: models/InteractionModule.py, DynamicInteraction.py, Cells.py, Router.py,
 * SelfAttention.py, Refinement.py, XModules.py).  The reference has NO FFI layer of its
 * own -- its boundary is the Python nn.Module API -- so every entry
 * point below cites the reference *Python* code whose arithmetic it replaces.  The host
 * mirror in d2r_b200/interaction/ binds these through ctypes.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated;
 *  - the caller owns every buffer (inputs, outputs, saved activations, workspace); the
 *    library never allocates, frees or retains device memory across calls;
 *  - every launch goes to the `stream` argument (a cudaStream_t passed as void*);
 *  - return value: 0 = OK, negative = error (see d2r_status); never throws, never exits.
 *    d2r_last_error() returns a thread-local human readable message for the last failure;
 *  - dtype enums select the arithmetic path: D2R_BF16 operands run on tcgen05 tensor
 *    cores (fp32 accumulation in TMEM), D2R_F32 operands run the fp32 CUDA-core path.
 *    There is no CPU fallback.
 */
#ifndef D2R_B200_H_
#define D2R_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D2R_B200_ABI_VERSION 1

typedef enum { D2R_OK = 0, D2R_ERR_ARG = -1, D2R_ERR_UNSUPPORTED = -2, D2R_ERR_CUDA = -3 } d2r_status;
typedef enum { D2R_F32 = 0, D2R_BF16 = 1 } d2r_dtype;
typedef enum { D2R_ACT_NONE = 0, D2R_ACT_RELU = 1, D2R_ACT_TANH = 2 } d2r_act;
typedef enum {
  D2R_EPI_STD = 0,    /* c = act(alpha*acc + bias) + residual                          */
  D2R_EPI_SQDIFF = 1, /* d = residual - (alpha*acc + bias);  c2 = d;  c = d*d          */
  /* Attention scores never leave the SM (bf16 tensor-core path, n <= 256 = one N tile):       */
  D2R_EPI_SOFTMAX = 2,     /* c = softmax_row(alpha*acc)            SelfAttention.py:33-37,    */
                           /*                                       XModules.py:306-309        */
  D2R_EPI_SOFTMAX_BWD = 3  /* c = P * (alpha*acc - sum_row(alpha*acc * P)),  P = residual      */
} d2r_epilogue;

/* ---- library probes ------------------------------------------------------------------ */
int d2r_abi_version(void);
/* "sm_100a" -- the only architecture the kernels are compiled for. */
const char* d2r_build_arch(void);
const char* d2r_last_error(void);
/* Number of kernels this library has launched in the calling process (all threads). */
int64_t d2r_launch_count(void);

/* ---- (b) contractions ---------------------------------------------------------------
 * One batched GEMM with a fused epilogue; replaces every F.linear / torch.matmul / bmm on
 * the path: Router.py:14-16, SelfAttention.py:29-39,53, XModules.py:300-310,
 * Refinement.py:105-115,134-137, Cells.py:149-158,198-209,238-246, InteractionModule.py:53.
 *
 *   C[z] = epilogue( alpha * A[z] (m x k) * B[z]^T (n x k) )
 *
 * A is stored [m,k] (a_mn_major=0, k contiguous) or [k,m] (a_mn_major=1); B is stored [n,k]
 * (b_mn_major=0, the nn.Linear weight layout) or [k,n] (b_mn_major=1).  Batch index
 * z = zo*batch_inner + zi addresses X + zo*x_so + zi*x_si (elements).  With
 * dtype=D2R_BF16 the kernel is TMA -> shared memory -> tcgen05.mma -> TMEM -> epilogue and
 * needs: 16-byte aligned a/b, lda/ldb and batch strides multiples of 8 elements.
 */
typedef struct d2r_gemm_args {
  int32_t dtype;       /* d2r_dtype of A and B */
  int32_t c_dtype;     /* d2r_dtype of C (and c2) */
  int32_t m, n, k;
  int32_t a_mn_major, b_mn_major;
  int32_t batch, batch_inner;
  int32_t act;         /* d2r_act */
  int32_t epilogue;    /* d2r_epilogue */
  int32_t r_dtype;     /* d2r_dtype of residual */
  int32_t accumulate;  /* 1: C += result (fp32 C only, atomic) */
  int32_t split_k;     /* >1: split the k range over CTAs, fp32 C, atomic accumulation */
  float alpha;
  int32_t tile_n;      /* 0 = auto; 64/128/192/256 force the tensor-core N tile, 512 the CTA-pair and 1024 the 4-CTA cluster kernel (tuning knob) */
  int32_t act_cols;    /* >0: the activation applies to output columns < act_cols only */
  int32_t reserved0;
  const void* a;
  const void* b;
  void* c;
  void* c2;             /* second output (SQDIFF), same layout as C */
  const float* bias;    /* optional fp32 bias[n] */
  const void* residual; /* optional, indexed like C with its own ld/strides */
  int64_t lda, ldb, ldc, ldr;
  int64_t a_so, a_si, b_so, b_si, c_so, c_si, r_so, r_si;
  int64_t bias_sz;      /* bias stride per batch index z (0 = shared) */
} d2r_gemm_args;

int d2r_gemm(const d2r_gemm_args* args, void* stream);

/* Debug probe of the tensor-core GEMM (tools/gemm_stall.py; the one piece of process-global state besides
 * the launch counter): while `records` is non-null every d2r_gemm(D2R_BF16) launch writes one record of
 * `return value` int64 slots per CTA (grid <= 148 CTAs) into it -- SM-clock cycles its TMA producer spent
 * waiting for a free shared-memory stage, its MMA issuer waiting for operands / for a free TMEM accumulator,
 * and one epilogue warp waiting for an accumulator.  Pass NULL to switch it off (the default). */
int d2r_gemm_set_profile(int64_t* records);

/* ---- (b) fused attention (bf16 tensor-core path) ------------------------------------------
 * One kernel for  P = softmax_row(alpha * Q K^T),  O = P V (+ residual)  -- or, for the alignment cell
 * (Cells.py:149), d = residual - P V, out2 = d, out = d * d -- per (sample, head): the scores stay in TMEM,
 * the probabilities go from registers into the shared-memory operand of the second tcgen05.mma and are also
 * written to `p` for the backward.  Replaces the pair of d2r_gemm launches (EPI_SOFTMAX, then P V) for
 * SelfAttention.py:33-39 (16 heads, head dim 48), XModules.py:300-310 / Refinement.py:105-115 (single head,
 * dim 768, alpha = 100/sqrt(768)) and Cells.py:244-246 (single head, unscaled).
 * Operands are bf16, addressed X[b, row, h*hd + col] with row stride x_ld and rows*x_ld per sample.
 * Limits: Lc <= 128, hd = 16..64 or a multiple of 64, every ld a multiple of 8, 16-byte aligned pointers. */
typedef struct d2r_attn_args {
  int32_t B, heads, Lq, Lc, hd;
  int32_t mode;          /* 0: out = P V (+ residual);  1: squared difference (needs residual and out2) */
  float alpha;           /* softmax(alpha * q.k) */
  int32_t reserved0;
  const void* q;         /* [B, Lq, heads*hd] */
  const void* k;         /* [B, Lc, heads*hd] */
  const void* v;         /* [B, Lc, heads*hd] */
  int64_t q_ld, k_ld, v_ld;
  void* p;               /* written: [B, heads, Lq, p_ld] probabilities, p_ld >= Lc */
  int64_t p_ld;
  void* out;             /* written: [B, Lq, heads*hd] with row stride o_ld */
  void* out2;            /* written in mode 1, same layout as out */
  int64_t o_ld;
  const void* residual;  /* optional, [B, Lq, heads*hd] with row stride r_ld */
  int64_t r_ld;
} d2r_attn_args;
int d2r_attn_fwd(const d2r_attn_args* a, void* stream);
/* Backward of the attention core in ONE kernel (Lq <= 128 and Lc <= 128): with P from the forward,
 *   dP = sign * d_out V^T;  dS = alpha * P o (dP - rowsum(dP o P));  dq = dS K;  dk = dS^T Q;  dv = sign * P^T d_out.
 * `sign` multiplies d_out (folds the minus of the squared-difference epilogue).  dq / dk / dv are overwritten, each
 * addressed like its forward operand with its own row stride (they may be column slices of wider buffers). */
typedef struct d2r_attn_bwd_args {
  int32_t B, heads, Lq, Lc, hd;
  int32_t reserved0;
  float alpha, sign;
  const void* d_out;     /* [B, Lq, heads*hd], row stride do_ld */
  const void* p;         /* [B, heads, Lq, p_ld] */
  const void* q;
  const void* k;
  const void* v;
  int64_t do_ld, p_ld, q_ld, k_ld, v_ld;
  void* dq;
  void* dk;
  void* dv;
  int64_t dq_ld, dk_ld, dv_ld;
} d2r_attn_bwd_args;
int d2r_attn_bwd(const d2r_attn_bwd_args* a, void* stream);

/* Row softmax over the last dim, optional scale: y = softmax(scale * x).  x fp32 or bf16
 * [rows, cols] with row stride ldx; y bf16 or fp32 with row stride ldy.
 * SelfAttention.py:33-37, XModules.py:306-309, Cells.py:244-245.  */
int d2r_softmax_fwd(const void* x, int32_t x_dtype, int64_t ldx, void* y, int32_t y_dtype, int64_t ldy,
                    int64_t rows, int32_t cols, float scale, void* stream);
/* dx = scale * y * (dy - sum(dy*y)) */
int d2r_softmax_bwd(const void* y, int32_t y_dtype, int64_t ldy, const void* dy, int32_t dy_dtype, int64_t lddy,
                    void* dx, int32_t dx_dtype, int64_t lddx, int64_t rows, int32_t cols, float scale,
                    void* stream);

/* ---- (a) router ------------------------------------------------------------------------
 * Router.py:22-26: soft_g = relu(tanh(W2 relu(W1 mean_L(x) + b1) + b2)).
 * d2r_pool_mean: pooled[g, b, :] = mean over L of x_g[b, :, :] for `groups` inputs given as
 * a device-visible pointer table passed BY VALUE (<= 8 groups).  The hidden layer is a
 * d2r_gemm; d2r_router_head applies W2/b2/tanh/relu for all K cells of a layer, then the
 * cross-cell normalisation and gate of DynamicInteraction.py:50-52 (final=0) or the raw
 * probabilities + per-cell gate of :104-117 (final=1).
 */
typedef struct d2r_ptr8 { const void* p[8]; } d2r_ptr8;
int d2r_pool_mean(d2r_ptr8 x, int32_t groups, int32_t x_dtype, int64_t B, int64_t L, int64_t D,
                  float* pooled /* [groups,B,D] */, void* stream);
int d2r_pool_mean_bwd(const float* d_pooled /* [B,D] */, int64_t B, int64_t L, int64_t D, void* dx,
                      int32_t dx_dtype, int32_t accumulate, void* stream);
/* hid: [K,B,H] fp32 (post-ReLU); w2: table of K pointers to [n_out,H]; b2: K pointers to [n_out].
 * raw[b,i,j] = relu(tanh(.)) with i = out path, j = cell; norm[b,i,j] = raw/(sum_j raw + 1e-8)
 * (final: norm = raw); gate[b,i] = sum_j raw < 1e-4 (final: gate[b,j] = raw[b,0,j] < 1e-4/K). */
int d2r_router_head_fwd(const float* hid, d2r_ptr8 w2, d2r_ptr8 b2, int32_t K, int32_t n_out, int64_t B,
                        int32_t H, int32_t final_layer, float* raw, float* norm, float* gate, void* stream);
/* d_norm [B,n_out,K] -> d_hid [K,B,H] (pre-ReLU mask applied with hid>0), dW2/db2 accumulated
 * into fp32 tables. */
int d2r_router_head_bwd(const float* d_norm, const float* raw, const float* hid, d2r_ptr8 w2, int32_t K,
                        int32_t n_out, int64_t B, int32_t H, int32_t final_layer, float* d_hid,
                        float* d_logit /* [B,n_out,K] scratch/out */, d2r_ptr8 d_w2, d2r_ptr8 d_b2, void* stream);

/* ---- (c) aggregation epilogue ----------------------------------------------------------
 * DynamicInteraction.py:55-67 (final=0) and :104-117 (final=1), fused with relu() of the RIC
 * cell (Cells.py:36-40) and with the mean-pool that feeds the next layer's routers.
 *   full[j]  : [B,L,D] tensors (cell 0 is the RAW RIC input, relu applied in-kernel)
 *   bvec[j]  : [B,D] vectors for broadcast cells (GLAC, GESC), NULL otherwise
 *   inputs[j]: final layer only: the layer inputs ref_wrd[j] for the gated skip
 *   P [B,n_out,K], gate [B,n_out] (final: [B,K])
 *   out[i]   : n_out tensors [B,L,D];  pooled: optional [n_out,B,D] mean over L of out[i]
 */
typedef struct d2r_agg_args {
  int32_t K, n_out, final_layer, dtype /* of full/inputs/out */;
  int64_t B, L, D;
  d2r_ptr8 full, bvec, inputs;
  d2r_ptr8 out;         /* written */
  const float* P;
  const float* gate;
  float* pooled;        /* optional */
} d2r_agg_args;
int d2r_aggregate_fwd(const d2r_agg_args* a, void* stream);
/* backward: d_out[i] (+ optional d_pooled [n_out,B,D]) -> d_full[j] (cell 0: gradient w.r.t. the
 * raw RIC input), d_bvec[j] [B,D], dP [B,n_out,K].  The gated-skip gradient of the final layer's
 * inputs j >= 1 is produced by d2r_gate_skip_bwd. */
typedef struct d2r_agg_bwd_args {
  d2r_agg_args fwd;     /* same tensors as forward (out[] unused) */
  d2r_ptr8 d_out;       /* n_out gradients [B,L,D] (dtype fwd.dtype) */
  const float* d_pooled;
  d2r_ptr8 d_full, d_bvec; /* written; d_bvec fp32 [B,D] */
  float* dP;             /* [B,n_out,K], overwritten */
} d2r_agg_bwd_args;
int d2r_aggregate_bwd(const d2r_agg_bwd_args* a, void* stream);
/* Final layer (DynamicInteraction.py:108-111): dx[j][b] (+)= gate[b,j] / S[b] * d_out[b] for j = 1..K-1,
 * S = sum_j (gate + P).  Bit j of accumulate_mask set: add into an existing gradient and touch only the
 * samples whose gate is set (p_j < 1e-4/K: essentially never); bit clear: overwrite (zero-fills the
 * un-gated samples).  NULL dx[j] entries are skipped. */
int d2r_gate_skip_bwd(const void* d_out, const float* P, const float* gate, d2r_ptr8 dx, int32_t K, int64_t B,
                      int64_t L, int64_t D, int32_t dtype, int32_t accumulate_mask, void* stream);

/* ---- small fused cell ops (elementwise / row reductions) -------------------------------
 * All take element counts and dtypes explicitly; `rows x cols` row-major with row stride =
 * cols unless an ld is given. */
/* y = x (cast), used for the per-step bf16 staging of fp32 parameters */
int d2r_cast(const void* x, int32_t x_dtype, void* y, int32_t y_dtype, int64_t n, void* stream);
/* dz = dy * act'(y); optional db[cols] += column sums of dz (fp32).  dz may alias dy or be NULL
 * (act none: only the bias gradient is produced). */
int d2r_bias_act_bwd(const void* dy, const void* y, int32_t dtype, int32_t act, void* dz, float* db,
                     int64_t rows, int32_t cols, int64_t ld, void* stream);
/* l2norm rows (XModules.py:14-18): y = x / (sqrt(sum x^2) + 1e-8); rnorm[rows] saved */
int d2r_l2norm_fwd(const void* x, int32_t dtype, void* y, float* rnorm, int64_t rows, int32_t cols,
                   void* stream);
int d2r_l2norm_bwd(const void* y, const void* dy, int32_t dtype, const float* rnorm, void* dx, int64_t rows,
                   int32_t cols, void* stream);
/* FiLM (Refinement.py:134-136): m = x * s + t, s = tanh-activated scale (already tanh'd), t shift;
 * st is [rows, 2*cols] = [s | t]. */
int d2r_film_fwd(const void* x, const void* st, int32_t dtype, void* m, int64_t rows, int32_t cols,
                 void* stream);
/* dx = dm*s (+ add, optional), d_st = [dm*x*(1-s^2) | dm] */
int d2r_film_bwd(const void* dm, const void* x, const void* st, const void* add, int32_t dtype, void* dx,
                 void* d_st, int64_t rows, int32_t cols, void* stream);
/* generic y = a*x + b*z elementwise (grad accumulation, residual joins) */
int d2r_axpby(const void* x, const void* z, int32_t dtype, float a, float b, void* y, int64_t n, void* stream);
/* squared difference backward (Cells.py:149): sq = d*d with d = x - c saved;  g = 2 d dsq
 * (dc = -g); gx = g (+ add, optional) is the gradient of x, written only when gx != NULL */
int d2r_sqdiff_bwd(const void* dsq, const void* d, const void* add, int32_t dtype, void* g, void* gx, int64_t n,
                   void* stream);
/* y = alpha * x * z elementwise */
int d2r_mul(const void* x, const void* z, int32_t dtype, float alpha, void* y, int64_t n, void* stream);

/* Attention filtration (XModules.py:380-384) over sim_emb = [global ; local]:
 *   logit[b,l] = w . S[b,l,:] + bias;  BN1d(1) over all B*(L+1) scalars (training: batch
 *   statistics + running-stat update with momentum 0.1; eval: running statistics);
 *   a = l1norm(sigmoid(.));  out[b,:] = l2norm(sum_l a[b,l] S[b,l,:]).
 * sg: [B,D] (global row, l = 0), sl: [B,L,D] (local rows).  saved: logits [B,L+1], attn [B,L+1],
 * stats[2] = {mean, invstd}, rnorm [B] = 1/(||saf|| + 1e-8). */
typedef struct d2r_saf_args {
  int32_t dtype, training;
  int64_t B, L, D;
  const void* sg; const void* sl;
  const float* w; const float* bias;       /* attn_sim_w.weight [D], .bias [1] */
  const float* bn_w; const float* bn_b;    /* [1] each */
  float* running_mean; float* running_var; int64_t* num_batches_tracked;  /* updated when training */
  float* logits; float* attn; float* stats; float* rnorm;
  float* out;                               /* [B,D] fp32 */
} d2r_saf_args;
int d2r_saf_fwd(const d2r_saf_args* a, void* stream);
typedef struct d2r_saf_bwd_args {
  d2r_saf_args fwd;
  const float* d_out;      /* [B,D] */
  void* d_sg; void* d_sl;  /* same dtype as inputs */
  float* d_w; float* d_bias; float* d_bn_w; float* d_bn_b;   /* accumulated (+=) */
  float* scratch;          /* fp32 workspace, >= B*D + B*(L+1) + 2 elements */
} d2r_saf_bwd_args;
int d2r_saf_bwd(const d2r_saf_bwd_args* a, void* stream);

/* GESC gate (Cells.py:205-209): g = softmax_D(gl); out = g*t + (1-g)*i   (all [B,D] fp32) */
int d2r_gate_fuse_fwd(const float* gl, const float* t, const float* i, float* g, float* out, int64_t B,
                      int32_t D, void* stream);
int d2r_gate_fuse_bwd(const float* d_out, const float* g, const float* t, const float* i, float* d_gl,
                      float* d_t, float* d_i, int64_t B, int32_t D, void* stream);

/* ---- (f, next) path-similarity loss ------------------------------------------------------
 * XModules.py:32-41 js_div(p, q), called on (sim_paths, sim_text) and (Reversed_sim_paths, sim_vision),
 * modeling_unimo.py:849.  p, q: fp32 [rows, cols] (row stride = cols).  With get_softmax != 0:
 *   P = softmax_row(p), Q = softmax_row(q), M = (P + Q) / 2,
 *   loss = ( sum P (log P - log M) + sum Q (log Q - log M) ) / (2 * rows)     (KLDivLoss 'batchmean')
 * get_softmax == 0: p and q already are the probabilities.  `loss` is ONE fp32 value, ACCUMULATED (+=): zero
 * it first.  The backward writes dp, dq = d_loss[0] * d loss / d p, d q (d_loss: device scalar). */
int d2r_js_div_fwd(const float* p, const float* q, int64_t rows, int32_t cols, int32_t get_softmax, float* loss,
                   void* stream);
int d2r_js_div_bwd(const float* p, const float* q, int64_t rows, int32_t cols, int32_t get_softmax,
                   const float* d_loss, float* dp, float* dq, void* stream);

/* Block bilinear fusion core (XModules.py:531-545, pos_norm='before_cat'), applied to the two pooled branch
 * outputs right after the stack (modeling_unimo.py:871-884).  Per sample b and chunk c (size S, rank R):
 *   r[b,c,s] = sum_{k<R} m0[b,c,k*S+s] * m1[b,c,k*S+s]
 *   zs = sign(r) sqrt(|r|);   z[b,c,:] = zs / max(||zs||_2, 1e-12)
 * m0, m1: [B, C*R*S] (dtype), the outputs of the chunk-wise merge linears (one batched d2r_gemm each);
 * z: [B, C*S] (dtype; the operand of linear_out); r [B, C*S] and inv_norm [B, C] (fp32) are kept for the
 * backward, which maps dz [B, C*S] to dm0, dm1 [B, C*R*S].  S <= 128. */
int d2r_block_merge_fwd(const void* m0, const void* m1, int32_t dtype, int64_t B, int32_t C, int32_t R, int32_t S,
                        void* z, float* r, float* inv_norm, void* stream);
int d2r_block_merge_bwd(const void* dz, const void* m0, const void* m1, const float* r, const float* inv_norm,
                        int32_t dtype, int64_t B, int32_t C, int32_t R, int32_t S, void* dm0, void* dm1,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* D2R_B200_H_ */
