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

}  // namespace b200
