/*
 * b200_hotpath.h -- C ABI of the B200-native training-step hot path.
 *
 * The reference (arthurcolle/mlx-cuda-distributed-pretraining) has no native code and no FFI: its
 * hot path is Python over mlx.core ops.  This header is therefore the boundary a maintainer would
 * bind *instead of* those mlx.core call sites; each entry point cites the reference lines whose
 * arithmetic it replaces.  All functions:
 *   - take raw device pointers + explicit sizes/strides + a cudaStream_t (passed as void*),
 *   - never allocate: scratch comes from the caller via *_workspace_bytes() queries,
 *   - return 0 on success, a negative b200 error code otherwise (message: b200_last_error()),
 *   - are asynchronous on `stream` and re-entrant across streams.
 * Element types: "bf16" = __nv_bfloat16, "f32" = float.  Matrices are row-major.
 */
#ifndef B200_HOTPATH_H_
#define B200_HOTPATH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_ARG (-1)
#define B200_ERR_CUDA (-2)
#define B200_ERR_DEVICE (-3)
#define B200_ERR_WORKSPACE (-4)

/* ---- library / device ---------------------------------------------------------------------- */
int b200_version(void);                 /* ABI version (monotonic integer) */
const char* b200_last_error(void);      /* thread-local message of the last failing call */
int b200_device_ok(void);               /* 1 if the current device is compute capability 10.x */
unsigned long long b200_launch_count(void); /* kernels launched by this library so far (process-wide) */
/* TMA descriptors are cached per (address, shape, strides, box): calls answered from the cache / encoded by the driver */
void b200_tensor_map_cache_stats(unsigned long long* hits, unsigned long long* misses);

/* ---- dense contraction engine (tcgen05 + TMA) ----------------------------------------------
 * D[b] = alpha*alpha_vec[b] * op(A[b]) op(B[b]) + beta*beta_vec[b] * C[b],  b in [0,batch)
 * replaces: `@` / mx.matmul in optimizers/muon.py:76-78, optimizers/shampoo.py:121,250,254,287-290
 * a_mn=0: A is [M,K] row-major (K-major);  a_mn=1: A is stored [K,M] row-major (MN-major)
 * b_mn=0: B is [N,K] row-major (K-major);  b_mn=1: B is stored [K,N] row-major (MN-major)
 * A,B bf16; C,D bf16 (out_f32=0) or f32 (out_f32=1); accumulate fp32 in TMEM.
 * alpha_vec/beta_vec: optional per-batch device scalars (NULL = 1).
 * force_bn: 0 auto | 128 | 256 tile width | 1256 = 256-wide tiles and the caller asserts D == D^T
 * (M == N, bf16 out): only tiles on/above the diagonal are computed, the rest mirror-written.
 * lda/ldb/ldd/ldc multiples of 8 (bf16) / 4 (f32 C,D).  M, K arbitrary (TMA zero-fills); N arbitrary too if
 * ldd/ldc >= round_up(N, 16 bytes): the padded tail columns of D are then written with alpha*0 + beta*C.
 */
int b200_gemm_bf16(int a_mn, int b_mn, int M, int N, int K, int batch,
                   const void* A, long long lda, long long strideA,
                   const void* B, long long ldb, long long strideB,
                   const void* C, long long ldc, long long strideC,
                   void* D, long long ldd, long long strideD,
                   int out_f32, float alpha, float beta,
                   const float* alpha_vec, const float* beta_vec, int force_bn, void* stream);

/* ---- Muon (optimizers/muon.py) ---------------------------------------------------------------
 * b200_newton_schulz: Muon.zeropower_via_newtonschulz5 (muon.py:54-83) on a batch of same-shape
 * matrices.  x_in: bf16 [batch,rows,cols], *un-normalised*; inv_norm[b] = 1/(||x_b||_F+eps) and
 * inv_norm_sq[b] = inv_norm[b]^2 (from b200_muon_momentum + b200_ns_scales) fold muon.py:72-73
 * into the first iteration.  x_out: bf16 [batch,rows,cols], must not alias x_in.
 */
size_t b200_newton_schulz_workspace_bytes(int batch, int rows, int cols, int steps);
int b200_newton_schulz(const void* x_in, void* x_out, int batch, int rows, int cols, int steps,
                       float a, float b, float c, const float* inv_norm, const float* inv_norm_sq,
                       void* workspace, size_t workspace_bytes, void* stream);
/* Same chain for the owner-computes data-parallel mode (SURVEY 8e "fused mode"; the reference's size-balanced
 * ownership idea: modal/modal_cuda_utils.py:468-490).  peer_out[i] (HOST array of n_peers <= 7 device
 * pointers) are peer-mapped addresses of the SAME x_out slot in the other ranks' buffers (NVLink peer
 * memory, e.g. torch symmetric memory).  The last X' = aX + BX GEMM writes its tiles to x_out and to every
 * peer from its epilogue -- GEMM and all-gather are one kernel; the caller only needs a cross-rank barrier
 * before reading the gathered buffer. */
int b200_newton_schulz_allgather(const void* x_in, void* x_out, int batch, int rows, int cols, int steps,
                                 float a, float b, float c, const float* inv_norm,
                                 const float* inv_norm_sq, void* workspace, size_t workspace_bytes,
                                 const void* const* peer_out, int n_peers, void* stream);

/* Scratch of the grid reductions below (b200_muon_momentum, b200_sumsq): per-block partials + arrival
 * counters.  The reductions are DETERMINISTIC (fixed summation order, no float atomics), so data-parallel
 * replicas that hold identical gradients compute bit-identical norms and stay bit-identical. */
size_t b200_reduce_workspace_bytes(int batch);
/* Whole-model chain: all shape groups of the optimizer advance through the iteration together and every stage
 * (A = X X^T | B = bA + cAA | X' = aX + BX) is ONE grouped tcgen05 launch over all groups -- 3 launches per
 * iteration instead of 3 per group (muon.py:91 loops over parameters).  Per group the arguments mean what they
 * mean in b200_newton_schulz / _allgather; n_groups <= 6 takes the grouped path (each group min(rows, cols) > 128),
 * anything else falls back to per-group chains.  Results are identical to per-group calls up to the fp32
 * summation order of split-K partials. */
typedef struct b200_ns_group {
  const void* x_in;            /* bf16 [batch, rows, cols], un-normalised */
  void* x_out;                 /* bf16 [batch, rows, cols] */
  int batch, rows, cols;
  const float* inv_norm;       /* [batch] 1/(||x||_F + eps) */
  const float* inv_norm_sq;    /* [batch] its square */
  const void* const* peer_out; /* HOST array of n_peers peer-mapped copies of x_out (NULL when n_peers == 0) */
  int n_peers;
} b200_ns_group;
size_t b200_newton_schulz_multi_workspace_bytes(const b200_ns_group* groups, int n_groups, int steps);
int b200_newton_schulz_multi(const b200_ns_group* groups, int n_groups, int steps, float a, float b, float c,
                             void* workspace, size_t workspace_bytes, void* stream);

/* muon.py:101,105: buf = (1-mu)*g*gscale + mu*buf ; u = nesterov ? g*gscale + mu*buf : buf.
 * g: [batch,numel] bf16 (g_is_bf16=1) or f32; buf f32; u bf16; sumsq[b] = sum(u_b^2) (overwritten). */
int b200_muon_momentum(const void* g, int g_is_bf16, float* buf, void* u_bf16, float* sumsq,
                       long long numel, int batch, float mu, int nesterov, float gscale,
                       void* workspace, size_t workspace_bytes, void* stream);
/* muon.py:72-73: inv_norm = 1/(sqrt(sumsq)+eps), inv_norm_sq = inv_norm^2 */
int b200_ns_scales(const float* sumsq, float* inv_norm, float* inv_norm_sq, int batch, float eps,
                   void* stream);
/* muon.py:111-114 (and the in-place apply the reference forgot):
 * p32 += s*x ; p16 = bf16(p32) if p16 != NULL.  x bf16 (x_is_bf16=1) or f32. */
int b200_axpy_update(float* p32, void* p16, const void* x, int x_is_bf16, long long n, float s,
                     void* stream);
/* muon.py:123-138 non-2-D fallback: momentum as above, p -= lr*u */
int b200_sgd_momentum(float* p32, void* p16, const void* g, int g_is_bf16, float* buf,
                      long long n, float mu, int nesterov, float lr, float gscale, void* stream);

/* ---- AdamW (core/training.py:821 -> mlx.optimizers.AdamW; enhanced_optimizers.py:157-184) ----
 * p *= 1-lr*wd ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
 * p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)        (bc1=bc2=1: no bias correction) */
int b200_adamw(float* p32, void* p16, const void* g, int g_is_bf16, float* m, float* v,
               long long n, float lr, float b1, float b2, float eps, float wd, float bc1,
               float bc2, float gscale, void* stream);
/* Adam step direction without applying it (Shampoo grafting, optimizers/shampoo.py:162-167,326) */
int b200_adam_direction(float* d, const void* g, int g_is_bf16, float* m, float* v, long long n,
                        float lr, float b1, float b2, float eps, float bc1, float bc2,
                        float gscale, void* stream);

/* ---- gradient post-processing (core/training.py:1664-1666,1671-1680) --------------------------
 * acc = (init?0:acc) + clamp(g,-clip,clip)*scale ; clip<=0 disables the clamp */
int b200_clip_accum(const void* g, int g_is_bf16, float* acc, long long n, float clip,
                    float scale, int init, void* stream);
/* out[b] (+)= sum(x[b,:]^2)   (optimizers/shampoo.py:300-301 Frobenius norms); workspace from
 * b200_reduce_workspace_bytes(batch) */
int b200_sumsq(const void* x, int x_is_bf16, float* out, long long numel, int batch,
               int zero_first, void* workspace, size_t workspace_bytes, void* stream);
/* hi = bf16((src + diag_add*I)*scale), lo = bf16(that - hi) (lo may be NULL); strided [rows,cols] */
int b200_split_bf16(const float* src, long long ld_src, void* hi, void* lo, long long ld_dst,
                    int rows, int cols, float scale, float diag_add, void* stream);

/* ---- Shampoo elementwise pieces (optimizers/shampoo.py:351-359, 365-373) ----------------------
 * ema_split: m = beta*m + (1-beta)*g*gscale ; mhat = m*inv_bc -> out32 = out_scale*mhat (f32),
 * hi/lo = bf16 split of mhat */
int b200_ema_split(const void* g, int g_is_bf16, float* m, float* out32, void* hi, void* lo,
                   long long n, float beta, float gscale, float inv_bc, float out_scale,
                   void* stream);
/* p[b] = p[b]*decay + coef[b]*pre[b] + coef_d[b]*d[b] ; p16 = bf16(p) (grafting + decoupled wd) */
int b200_graft_update(float* p32, void* p16, const float* pre, const float* d, long long numel,
                      int batch, const float* coef, const float* coef_d, float decay, void* stream);

/* ---- Shampoo's Kronecker-factor path (optimizers/shampoo.py), batched over same-shape parameters --------
 * Factor matrices (statistics L/R, preconditioners P) are f32 [batch, kp, kp], kp = round_up(k, 8), zero
 * padded; their bf16 hi/lo splits have the same layout.  fp32 operands enter the tensor cores as bf16 hi+lo
 * pairs (three accumulating GEMMs); a NULL lo pointer means "plain bf16 operand".
 *
 * b200_shampoo_stats -- Shampoo._update_statistics (shampoo.py:229-255):
 *   L = beta2*L + weight * G[:k1,:k2] G[:k1,:k2]^T ;  R = beta2*R + weight * G[:k1,:k2]^T G[:k1,:k2]
 *   g_hi (/g_lo): bf16 [batch][rows >= k1][ldg >= k2] (the parameter-shaped gradient, read in place);
 *   weight = (1-beta2) * gscale^2 when the stored gradient still carries a 1/gscale factor. */
int b200_shampoo_stats(const void* g_hi, const void* g_lo, long long ldg, long long strideG, float* L,
                       float* R, int batch, int k1, int k2, float beta2, float weight, void* stream);
/* b200_shampoo_root -- MatrixSqrt.matrix_inverse_pth_root (shampoo.py:88-126), reproduced literally:
 *   Mt = M + eps*I ; Z = Mt/tr(Mt) ; iters x { Z = Z @ (I + Z/p) } ; P = Z * tr(Mt)^(-1/p^2)
 *   M, P: f32 [batch,kp,kp] (distinct); P_hi/P_lo (nullable): bf16 split of P for b200_shampoo_precond. */
size_t b200_shampoo_root_workspace_bytes(int batch, int k);
int b200_shampoo_root(const float* M, float* P, void* P_hi, void* P_lo, int batch, int k, float p,
                      float eps, int iters, void* workspace, size_t workspace_bytes, void* stream);
/* b200_shampoo_precond -- Shampoo._apply_preconditioners (shampoo.py:257-295):
 *   out[:k1,:k2] = alpha * PL @ m[:k1,:k2] @ PR   (elements of `out` outside the block are left alone)
 *   m_hi/m_lo: bf16 split of the (bias-corrected) momentum, [batch][rows][ldm]; out: f32 [batch][rows][ldo];
 *   k2 % 8 == 0, k1 arbitrary. */
size_t b200_shampoo_precond_workspace_bytes(int batch, int k1, int k2);
int b200_shampoo_precond(const void* PL_hi, const void* PL_lo, const void* PR_hi, const void* PR_lo,
                         const void* m_hi, const void* m_lo, long long ldm, long long strideM, float* out,
                         long long ldo, long long strideO, int batch, int k1, int k2, float alpha,
                         void* workspace, size_t workspace_bytes, void* stream);
/* b200_shampoo_graft -- Shampoo._apply_grafting + the parameter write (shampoo.py:297-312,365-373):
 *   sn = ||upd_b||_F, gn = ||graft_b||_F (f32, overflow to inf like mx.linalg.norm);
 *   step = sn==0 ? graft : (gn==0 ? upd : upd*gn/sn) ;  p = p*decay + step ; p16 = bf16(p) if non-NULL.
 *   upd, graft, p32: f32 [batch, numel]. Deterministic reductions. */
size_t b200_shampoo_graft_workspace_bytes(int batch);
int b200_shampoo_graft(float* p32, void* p16, const float* upd, const float* graft, long long numel,
                       int batch, float decay, void* workspace, size_t workspace_bytes, void* stream);

/* ---- RMSNorm (arch/llama.py:50-56) and RoPE (arch/llama_standard.py:74-75,117-127) ------------ */
int b200_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int rows, int H,
                     float eps, int is_bf16, void* stream);
size_t b200_rmsnorm_bwd_workspace_bytes(int rows, int H);   /* per-CTA dW partials */
int b200_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx,
                     float* dw_f32 /* NULL: deferred, see below */, int rows, int H, int is_bf16, void* workspace,
                     size_t workspace_bytes, void* stream);
/* Deferred weight gradient: (add_)rmsnorm_bwd called with dw_f32 == NULL leaves its per-CTA partial sums
 * [b200_rmsnorm_bwd_partial_rows(rows, H)][H] (fp32) in the workspace; b200_rmsnorm_dw_reduce then folds any number of
 * such workspaces (one job per norm, all of width H) in one launch per 32 jobs, writing or accumulating into each
 * weight's gradient in its storage type -- the backward of an L-layer model reduces its 2L+1 norm weights once, after
 * the last layer, instead of with four small launches per norm. */
typedef struct b200_dw_job {
  const float* partials; /* workspace of the matching (add_)rmsnorm_bwd call */
  void* dw;              /* [H] bf16 or fp32 */
  int n_partials;        /* b200_rmsnorm_bwd_partial_rows(rows, H) of that call */
  int dw_is_bf16;
  int accumulate;        /* 1: dw += sum (autograd semantics), 0: dw = sum */
  int reserved;
} b200_dw_job;
int b200_rmsnorm_bwd_partial_rows(int rows, int H);
int b200_rmsnorm_dw_reduce(const b200_dw_job* jobs, int n_jobs, int H, void* stream);
/* Residual add fused with the norm that follows it (arch/llama.py:316-319: h = x + sublayer(...), then
 * the next norm reads h):  sum_out = x + delta (rounded to the storage type), y = rmsnorm(sum_out) * w.
 * Backward: dx = dres + rmsnorm_bwd(dy) where dres (nullable) is the gradient that reached sum_out
 * directly; dx is the gradient of both x and delta.  Same workspace as b200_rmsnorm_bwd. */
int b200_add_rmsnorm_fwd(const void* x, const void* delta, const void* w, void* sum_out, void* y,
                         float* rstd, int rows, int H, float eps, int is_bf16, void* stream);
int b200_add_rmsnorm_bwd(const void* dy, const void* dres, const void* x, const void* w,
                         const float* rstd, void* dx, float* dw_f32, int rows, int H, int is_bf16,
                         void* workspace, size_t workspace_bytes, void* stream);
/* x,y: [B,S,NH,D]; cos_t,sin_t: f32 [S,D/2]; backward=1 applies the inverse rotation */
int b200_rope(const void* x, void* y, const float* cos_t, const float* sin_t, int B, int S, int NH,
              int D, int backward, int is_bf16, void* stream);

/* ---- MLP block on the in-tree GEMM engine with the GLU in the epilogue ----------------
 * arch/llama.py:142-151: down( gate(x) * sigmoid(up(x)) * 2 ).  gate_proj.weight and up_proj.weight must be
 * adjacent in memory (W2 = [Wg ; Wu], [2I, K] row-major; flat.ParamStore lays them out that way).
 * b200_mlp_gateup_glu_fwd:  gu[M, 2I] = x[M, K] W2^T  (columns [0, I) = gate(x), [I, 2I) = up(x)),
 *                           y[M, I] = gate * sigmoid(up) * 2 from the fp32 accumulators -- one tcgen05 kernel,
 *                           no separate activation pass.  K % 8 == 0, I % 128 == 0.
 * b200_mlp_down_glu_bwd:    d = dy[M, H] Wd (dgrad of down_proj, Wd = down_proj.weight [H, I]) never leaves the
 *                           chip: dgu[M, 2I] = [ d * sigmoid(u) * 2 | d * g * sigmoid(u) * (1 - sigmoid(u)) * 2 ]
 *                           with g, u read from the saved gu.  [dg | du] is one matrix, so the gate/up dgrad and
 *                           wgrad each run as a single GEMM against W2. */
int b200_mlp_gateup_glu_fwd(const void* x, const void* W2, void* gu, void* y, int M, int K, int I, void* stream);
int b200_mlp_down_glu_bwd(const void* dy, const void* Wd, const void* gu, void* dgu, int M, int H, int I,
                          void* stream);

/* ---- block-adjacent fused elementwise steps ------------------------------------
 * arch/llama.py:151: y = gate * sigmoid(up) * 2 and its adjoint; bf16, n % 8 == 0 */
int b200_glu_fwd(const void* gate, const void* up, void* y, long long n, void* stream);
int b200_glu_bwd(const void* dy, const void* gate, const void* up, void* dgate, void* dup, long long n,
                 void* stream);
/* Adjoint of the embedding gather (arch/llama.py:389): grad[tokens[r], :] += dh[r, :] for r in [0, rows);
 * dh bf16 [rows, H]; grad bf16 or f32 [V, H] (accumulated into: it may already hold the tied-logits wgrad);
 * fp32 accumulation per vocabulary row in the workspace (V*H floats), tokens outside [0, V) are ignored. */
size_t b200_embedding_bwd_workspace_bytes(int V, int H);
int b200_embedding_bwd(const void* dh, const long long* tokens, void* grad, int grad_is_bf16, long long rows,
                       int V, int H, void* workspace, size_t workspace_bytes, void* stream);
/* core/training.py:1226-1234: per-row cross entropy of bf16 logits [rows, ld] (first V columns valid),
 * computed in fp32: row_lse = logsumexp, row_loss = (target != pad_token) ? lse - logit[target] : 0.
 * ce_bwd overwrites logits with row_scale[row] * (softmax - onehot(target)) (0 in columns >= V and in
 * rows whose target is pad_token). */
int b200_ce_fwd(const void* logits, long long ld, const long long* targets, int rows, int V,
                long long pad_token, float* row_loss, float* row_lse, void* stream);
int b200_ce_bwd(void* logits, long long ld, const long long* targets, int rows, int V, long long pad_token,
                const float* row_lse, const float* row_scale, void* stream);

/* ---- fused causal / GQA attention (arch/flash_attention.py:97-156 + its autograd) --------------
 * q: bf16 [B,S,H,D], k,v: bf16 [B,S,Hk,D] (q head h uses kv head h/(H/Hk), flash_attention.py:102-120)
 * o: bf16 [B,S,H,D]; lse: f32 [B,H,S] (natural-log sum-exp of scaled, masked scores)
 * causal=1 reproduces the additive -inf strictly-upper mask of arch/llama.py:384-387. */
int b200_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int S,
                  int H, int Hk, int D, float scale, int causal, void* stream);
size_t b200_attn_bwd_workspace_bytes(int B, int S, int H, int Hk, int D);
int b200_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                  const float* lse, void* dq, void* dk, void* dv, int B, int S, int H, int Hk,
                  int D, float scale, int causal, void* workspace, size_t workspace_bytes,
                  void* stream);
/* Same, with dk / dv token rows of dkv_row_heads heads (>= Hk).  dk = base, dv = base + Hk*D,
 * dkv_row_heads = 2*Hk writes both into one [B,S,2*Hk*D] buffer, so the k/v projections' dgrad / wgrad
 * (arch/flash_attention.py:52-53 weights, adjacent in the flat store) each run as ONE GEMM. */
int b200_attn_bwd_strided(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                          const float* lse, void* dq, void* dk, void* dv, int B, int S, int H, int Hk,
                          int D, float scale, int causal, int dkv_row_heads, void* workspace,
                          size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_HOTPATH_H_ */
