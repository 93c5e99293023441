// Newton-Schulz orthogonalisation driver (reference: Muon.zeropower_via_newtonschulz5,
// optimizers/muon.py:54-83):   X0 = G/(||G||_F+eps);  repeat: A = X X^T, B = bA + cAA, X = aX + BX
//
// B200 formulation: every step is a batched tcgen05 GEMM with a fused epilogue
//   G1: A  = X X^T                (wide)   |  A = X^T X   (tall, both operands MN-major)
//   G2: B  = b*A + c*(A A)        (A symmetric -> A A^T, both operands K-major)
//   G3: X' = a*X + B X            (wide, X as MN-major B operand) | X' = a*X + X B (tall)
// so the reference's transpose for tall matrices (muon.py:68-70,80-81) never materialises, and
// the normalisation X0 = G/(||G||+eps) is folded into the first iteration's epilogue scalars
// (per-matrix 1/(norm+eps) vectors produced by the momentum kernel's fused sum of squares).
#include <stdlib.h>

#include "host.h"

namespace b200 {

int gemm_bf16(bool a_mn, bool b_mn, int M, int N, int K, int batch, const void* A, long long lda,
              long long strideA, const void* B, long long ldb, long long strideB, const void* C,
              long long ldc, long long strideC, void* D, long long ldd, long long strideD,
              bool out_f32, float alpha, float beta, const float* alpha_vec, const float* beta_vec,
              int force_bn, cudaStream_t stream);

int gemm_bf16_2cta(bool a_mn, bool b_mn, int M, int N, int K, int batch, const void* A, long long lda,
                   long long strideA, const void* B, long long ldb, long long strideB, const void* C,
                   long long ldc, long long strideC, void* D, long long ldd, long long strideD, bool out_f32,
                   float alpha, float beta, const float* alpha_vec, const float* beta_vec, int bn,
                   int symmetric, int k_splits, float* splitk_ws, const void* const* peer_D, int n_peers,
                   cudaStream_t stream);

static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// One shape group of the whole-model chain (mirrors b200_ns_group of the public header)
struct NsGroup {
  const void* x_in;
  void* x_out;
  int batch, rows, cols;
  const float* inv_norm;
  const float* inv_norm_sq;
  const void* const* peer_out;
  int n_peers;
};

// G1 of a lone big matrix (the 32003 x 1024 embedding: 10 symmetric tiles, K = 32003) would occupy a
// fraction of the CTA pairs for hundreds of microseconds.  Cut K so that every pair gets a work item.
static int g1_k_splits(int batch, int m, int k) {
  if (m <= 128 || (m & 7) != 0) return 1;
  const int t = (m + 255) / 256;
  const long long tiles = (long long)t * (t + 1) / 2 * batch;
  const int pairs = num_sms() / 2;
  if (tiles * 2 > pairs) return 1;
  const int num_kb = (k + 63) / 64;
  int s = (int)(pairs / tiles);
  if (s > num_kb / 16) s = num_kb / 16;
  return s < 2 ? 1 : s;
}

// A = X X^T and B = bA + cAA are symmetric: only tiles on or above the diagonal are computed and
// the rest mirror-written (force_bn code understood by gemm_bf16; ignored when it cannot apply)
static const int kSym = [] {
  const char* e = getenv("B200_NS_NOSYM");  // debug: full (non-symmetric) G1/G2 tiles
  return (e != nullptr && e[0] == '1') ? 0 : 1256;
}();

size_t newton_schulz_workspace_bytes(int batch, int rows, int cols, int steps) {
  const size_t m = rows < cols ? rows : cols;
  size_t bytes = 2 * align256((size_t)batch * m * m * 2);
  if (steps >= 2) bytes += align256((size_t)batch * rows * cols * 2);
  const int splits = g1_k_splits(batch, (int)m, rows < cols ? cols : rows);
  if (splits > 1) bytes += align256((size_t)splits * batch * m * m * 4);
  return bytes;
}

// peer_out / n_peers: peer-mapped copies of x_out on other GPUs (owner-computes Newton-Schulz under data
// parallelism).  The LAST iteration's X' = aX + BX GEMM then stores its tiles to x_out and to every peer
// from its epilogue (GEMM -> all-gather in one kernel); with n_peers == 0 this is the plain chain.
int newton_schulz(const void* x_in, void* x_out, int batch, int rows, int cols, int steps, float a,
                  float b, float c, const float* inv_norm, const float* inv_norm_sq, void* ws,
                  size_t ws_bytes, const void* const* peer_out, int n_peers, cudaStream_t stream) {
  B200_CHECK_ARG(n_peers >= 0 && n_peers <= 7 && (n_peers == 0 || peer_out != nullptr),
                 "newton_schulz: bad peer list (n_peers=%d)", n_peers);
  B200_CHECK_ARG(batch > 0 && rows > 0 && cols > 0 && steps > 0,
                 "newton_schulz: bad shape batch=%d rows=%d cols=%d steps=%d", batch, rows, cols,
                 steps);
  B200_CHECK_ARG(cols % 8 == 0, "newton_schulz: cols=%d must be a multiple of 8 (16-byte rows)",
                 cols);
  B200_CHECK_ARG(x_in != x_out, "newton_schulz: x_in and x_out must be distinct buffers");
  if (ws_bytes < newton_schulz_workspace_bytes(batch, rows, cols, steps)) {
    set_error("newton_schulz: workspace too small (%zu < %zu)", ws_bytes,
              newton_schulz_workspace_bytes(batch, rows, cols, steps));
    return B200_ERR_WORKSPACE;
  }
  const bool tall = rows > cols;
  const int m = tall ? cols : rows;
  const long long mm = (long long)m * m;
  const long long rc = (long long)rows * cols;
  char* w = reinterpret_cast<char*>(ws);
  void* Abuf = w;
  void* Bbuf = w + align256((size_t)batch * mm * 2);
  void* Xtmp = w + 2 * align256((size_t)batch * mm * 2);
  const int splits = g1_k_splits(batch, m, tall ? rows : cols);
  float* splitk_ws = reinterpret_cast<float*>(
      w + 2 * align256((size_t)batch * mm * 2) + (steps >= 2 ? align256((size_t)batch * rc * 2) : 0));

  const void* cur = x_in;
  for (int it = 0; it < steps; ++it) {
    void* nxt = ((steps - 1 - it) % 2 == 0) ? x_out : Xtmp;
    const float* s1 = (it == 0) ? inv_norm : nullptr;
    const float* s2 = (it == 0) ? inv_norm_sq : nullptr;
    int rc_;
    // G1
    if (splits > 1)
      rc_ = gemm_bf16_2cta(tall, tall, m, m, tall ? rows : cols, batch, cur, cols, rc, cur, cols, rc, nullptr,
                           0, 0, Abuf, m, mm, false, 1.0f, 0.0f, s2, nullptr, 256, 1, splits, splitk_ws,
                           nullptr, 0, stream);
    else if (!tall)
      rc_ = gemm_bf16(false, false, m, m, cols, batch, cur, cols, rc, cur, cols, rc, nullptr, 0, 0,
                      Abuf, m, mm, false, 1.0f, 0.0f, s2, nullptr, kSym, stream);
    else
      rc_ = gemm_bf16(true, true, m, m, rows, batch, cur, cols, rc, cur, cols, rc, nullptr, 0, 0,
                      Abuf, m, mm, false, 1.0f, 0.0f, s2, nullptr, kSym, stream);
    if (rc_) return rc_;
    // G2: B = b*A + c*A*A
    rc_ = gemm_bf16(false, false, m, m, m, batch, Abuf, m, mm, Abuf, m, mm, Abuf, m, mm, Bbuf, m,
                    mm, false, c, b, nullptr, nullptr, kSym, stream);
    if (rc_) return rc_;
    // G3 (the last one carries the peer stores)
    const bool last = it == steps - 1;
    const int g3_m = tall ? rows : m, g3_n = tall ? m : cols;
    if (last && n_peers > 0 && g3_m > 128) {
      // CTA-pair kernel directly: tile-N choice as in gemm_bf16
      const long long t256 = (long long)((g3_m + 255) / 256) * ((g3_n + 255) / 256) * batch;
      const int bn2 = (g3_n <= 128 || t256 < (num_sms() / 2) * 3 / 4) ? 128 : 256;
      if (!tall)
        rc_ = gemm_bf16_2cta(false, true, m, cols, m, batch, Bbuf, m, mm, cur, cols, rc, cur, cols, rc, nxt, cols,
                             rc, false, 1.0f, a, s1, s1, bn2, 0, 1, nullptr, peer_out, n_peers, stream);
      else
        rc_ = gemm_bf16_2cta(false, false, rows, m, m, batch, cur, cols, rc, Bbuf, m, mm, cur, cols, rc, nxt, cols,
                             rc, false, 1.0f, a, s1, s1, bn2, 0, 1, nullptr, peer_out, n_peers, stream);
    } else {
      if (!tall)
        rc_ = gemm_bf16(false, true, m, cols, m, batch, Bbuf, m, mm, cur, cols, rc, cur, cols, rc,
                        nxt, cols, rc, false, 1.0f, a, s1, s1, 0, stream);
      else
        rc_ = gemm_bf16(false, false, rows, m, m, batch, cur, cols, rc, Bbuf, m, mm, cur, cols, rc,
                        nxt, cols, rc, false, 1.0f, a, s1, s1, 0, stream);
      if (rc_ == 0 && last && n_peers > 0) {
        // small matrices run on the single-CTA kernel, which has no peer epilogue: copy the result out
        for (int pi = 0; pi < n_peers; ++pi)
          B200_CHECK_CUDA(cudaMemcpyAsync(const_cast<void*>(peer_out[pi]), nxt, (size_t)batch * rc * 2,
                                          cudaMemcpyDeviceToDevice, stream));
      }
    }
    if (rc_) return rc_;
    cur = nxt;
  }
  return B200_OK;
}


// ---------------------------------------------------------------------------------------------------------
// Whole-model chain: every shape group of the optimizer advances through the SAME iteration together, so each
// stage (A = X X^T | B = bA + cAA | X' = aX + BX) is ONE grouped launch over all groups (gemm_tc2.cu) instead of
// one launch per group: 3 launches per iteration (+ one split-K finalize for a lone tall matrix) instead of
// 3 x groups, and the per-launch wave tail is amortised over all groups' tiles (C2: 682 instead of 240 tiles
// for the B stage on 74 CTA pairs).  Falls back to the per-group chain when a group is too small for the
// CTA-pair kernel (min(rows, cols) <= 128) or there are more groups than one launch takes.
// ---------------------------------------------------------------------------------------------------------
static bool multi_grouped_ok(const NsGroup* g, int n) {
  static const bool off = [] {
    const char* e = getenv("B200_NS_GROUPED");
    return e != nullptr && e[0] == '0';
  }();
  if (off || n < 1 || n > 6 || num_sms() < 2) return false;
  for (int i = 0; i < n; ++i) {
    const int m = g[i].rows < g[i].cols ? g[i].rows : g[i].cols;
    if (m <= 128 || (m & 7) != 0 || (g[i].cols & 7) != 0) return false;
  }
  return true;
}

struct NsLayout {  // workspace slices of one group
  size_t a, b, xtmp, splitk, end;
  int splits;
};
static NsLayout ns_layout(const NsGroup& g, int steps, size_t base) {
  const size_t m = g.rows < g.cols ? g.rows : g.cols;
  NsLayout l;
  l.a = base;
  l.b = l.a + align256((size_t)g.batch * m * m * 2);
  l.xtmp = l.b + align256((size_t)g.batch * m * m * 2);
  l.splitk = l.xtmp + (steps >= 2 ? align256((size_t)g.batch * g.rows * g.cols * 2) : 0);
  l.splits = g1_k_splits(g.batch, (int)m, g.rows < g.cols ? g.cols : g.rows);
  l.end = l.splitk + (l.splits > 1 ? align256((size_t)l.splits * g.batch * m * m * 4) : 0);
  return l;
}

size_t newton_schulz_multi_workspace_bytes(const NsGroup* g, int n, int steps) {
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    const size_t grouped_end = ns_layout(g[i], steps, off).end;
    const size_t single_end = off + newton_schulz_workspace_bytes(g[i].batch, g[i].rows, g[i].cols, steps);
    off = grouped_end > single_end ? grouped_end : single_end;
  }
  return off;
}

int newton_schulz_multi(const NsGroup* g, int n, int steps, float a, float b, float c, void* ws, size_t ws_bytes,
                        cudaStream_t stream) {
  B200_CHECK_ARG(n >= 1 && g != nullptr && steps > 0, "newton_schulz_multi: bad arguments (n=%d steps=%d)", n, steps);
  for (int i = 0; i < n; ++i) {
    B200_CHECK_ARG(g[i].batch > 0 && g[i].rows > 0 && g[i].cols > 0 && g[i].cols % 8 == 0,
                   "newton_schulz_multi: group %d has a bad shape batch=%d rows=%d cols=%d", i, g[i].batch, g[i].rows,
                   g[i].cols);
    B200_CHECK_ARG(g[i].x_in != g[i].x_out, "newton_schulz_multi: group %d: x_in and x_out must be distinct", i);
    B200_CHECK_ARG(g[i].n_peers >= 0 && g[i].n_peers <= 7 && (g[i].n_peers == 0 || g[i].peer_out != nullptr),
                   "newton_schulz_multi: group %d: bad peer list", i);
  }
  if (ws_bytes < newton_schulz_multi_workspace_bytes(g, n, steps)) {
    set_error("newton_schulz_multi: workspace too small (%zu < %zu)", ws_bytes,
              newton_schulz_multi_workspace_bytes(g, n, steps));
    return B200_ERR_WORKSPACE;
  }
  char* w = reinterpret_cast<char*>(ws);
  if (!multi_grouped_ok(g, n)) {
    size_t off = 0;
    for (int i = 0; i < n; ++i) {
      const size_t grouped_end = ns_layout(g[i], steps, off).end;
      const size_t need = newton_schulz_workspace_bytes(g[i].batch, g[i].rows, g[i].cols, steps);
      int rc = newton_schulz(g[i].x_in, g[i].x_out, g[i].batch, g[i].rows, g[i].cols, steps, a, b, c, g[i].inv_norm,
                             g[i].inv_norm_sq, w + off, need, g[i].peer_out, g[i].n_peers, stream);
      if (rc) return rc;
      off = grouped_end > off + need ? grouped_end : off + need;
    }
    return B200_OK;
  }
  NsLayout lay[6];
  const void* cur[6];
  {
    size_t off = 0;
    for (int i = 0; i < n; ++i) {
      lay[i] = ns_layout(g[i], steps, off);
      const size_t single_end = off + newton_schulz_workspace_bytes(g[i].batch, g[i].rows, g[i].cols, steps);
      off = lay[i].end > single_end ? lay[i].end : single_end;
      cur[i] = g[i].x_in;
    }
  }
  GroupedGemm pr[6];
  for (int it = 0; it < steps; ++it) {
    const bool first = it == 0, last = it == steps - 1;
    // ---- G1: A = X X^T (wide) / X^T X (tall), symmetric; iteration 1 carries 1/(||G||+eps)^2 ----
    for (int i = 0; i < n; ++i) {
      const bool tall = g[i].rows > g[i].cols;
      const int m = tall ? g[i].cols : g[i].rows;
      const long long rc_ = (long long)g[i].rows * g[i].cols, mm = (long long)m * m;
      GroupedGemm& q = pr[i];
      q = GroupedGemm{};
      q.a_mn = tall;
      q.b_mn = tall;
      q.M = m;
      q.N = m;
      q.K = tall ? g[i].rows : g[i].cols;
      q.batch = g[i].batch;
      q.A = cur[i];
      q.lda = g[i].cols;
      q.strideA = rc_;
      q.B = cur[i];
      q.ldb = g[i].cols;
      q.strideB = rc_;
      q.D = w + lay[i].a;
      q.ldd = m;
      q.strideD = mm;
      q.alpha = 1.0f;
      q.beta = 0.0f;
      q.alpha_vec = first ? g[i].inv_norm_sq : nullptr;
      q.symmetric = kSym ? 1 : 0;
      q.k_splits = lay[i].splits;
      q.splitk_ws = reinterpret_cast<float*>(w + lay[i].splitk);
    }
    int rc = gemm_grouped_2cta(pr, n, stream);
    if (rc) return rc;
    // ---- G2: B = b A + c A A (A symmetric: A A^T with both operands K-major) ----
    for (int i = 0; i < n; ++i) {
      const int m = g[i].rows > g[i].cols ? g[i].cols : g[i].rows;
      const long long mm = (long long)m * m;
      GroupedGemm& q = pr[i];
      q = GroupedGemm{};
      q.M = q.N = q.K = m;
      q.batch = g[i].batch;
      q.A = q.B = q.C = w + lay[i].a;
      q.lda = q.ldb = q.ldc = q.ldd = m;
      q.strideA = q.strideB = q.strideC = q.strideD = mm;
      q.D = w + lay[i].b;
      q.alpha = c;
      q.beta = b;
      q.symmetric = kSym ? 1 : 0;
      q.k_splits = 1;
    }
    rc = gemm_grouped_2cta(pr, n, stream);
    if (rc) return rc;
    // ---- G3: X' = a X + B X (wide) / a X + X B (tall); the last one carries the peer stores ----
    for (int i = 0; i < n; ++i) {
      const bool tall = g[i].rows > g[i].cols;
      const int m = tall ? g[i].cols : g[i].rows;
      const long long rc_ = (long long)g[i].rows * g[i].cols, mm = (long long)m * m;
      void* nxt = ((steps - 1 - it) % 2 == 0) ? g[i].x_out : static_cast<void*>(w + lay[i].xtmp);
      GroupedGemm& q = pr[i];
      q = GroupedGemm{};
      if (!tall) {  // [m, cols] = B[m, m] (K-major) x X[m(K), cols] (MN-major)
        q.a_mn = false;
        q.b_mn = true;
        q.M = m;
        q.N = g[i].cols;
        q.K = m;
        q.A = w + lay[i].b;
        q.lda = m;
        q.strideA = mm;
        q.B = cur[i];
        q.ldb = g[i].cols;
        q.strideB = rc_;
      } else {  // [rows, m] = X[rows, m(K)] (K-major) x B[m, m] (symmetric: K-major read of B^T = B)
        q.a_mn = false;
        q.b_mn = false;
        q.M = g[i].rows;
        q.N = m;
        q.K = m;
        q.A = cur[i];
        q.lda = g[i].cols;
        q.strideA = rc_;
        q.B = w + lay[i].b;
        q.ldb = m;
        q.strideB = mm;
      }
      q.batch = g[i].batch;
      q.C = cur[i];
      q.ldc = g[i].cols;
      q.strideC = rc_;
      q.D = nxt;
      q.ldd = g[i].cols;
      q.strideD = rc_;
      q.alpha = 1.0f;
      q.beta = a;
      q.alpha_vec = first ? g[i].inv_norm : nullptr;
      q.beta_vec = first ? g[i].inv_norm : nullptr;
      q.k_splits = 1;
      if (last && g[i].n_peers > 0) {
        q.peer_D = g[i].peer_out;
        q.n_peers = g[i].n_peers;
      }
      cur[i] = nxt;
    }
    rc = gemm_grouped_2cta(pr, n, stream);
    if (rc) return rc;
  }
  return B200_OK;
}

}  // namespace b200
