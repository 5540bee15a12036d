// pgb_solve.cu -- the m x m penalised solve of one PIRLS iteration, fp64, on the device.
//
// Replaces pygam.py:783-805 (dense QR of W.B, SVD of [R; E], back-projection B, coef = B.Wz)
// and the edof / cov part of _estimate_model_statistics (pygam.py:1033-1045) by the
// algebraically equal normal-equation form
//     beta = (G + E^T E)^-1 r,  edof = tr(G (G + E^T E)^-1),  cov = A^-1 G A^-1,
// solved in an orthogonal basis Q = [Nhat | Z] whose first p columns span the analytic null
// space of the design (one vector per partition-of-unity term block).  In
// that basis the Nhat rows/columns of Q^T G Q and Nhat^T r are zero in exact arithmetic; the
// kernel sets them to exact zero, which removes the rounding noise that a plain Cholesky of
// G + E^T E amplifies by 1/sqrt(EPS) and that otherwise changes iteration counts.
//
// One CTA per system.  Small systems (two m x m matrices fit in shared memory) run entirely
// out of shared memory -- this is the grid-search path, one candidate per CTA; larger ones
// work in the caller's workspace (L2-resident).
#include <cuda_runtime.h>
#include <string.h>

#include <mutex>
#include <vector>
#include <math.h>

#include <algorithm>

#include "pgb_internal.h"

static const int kSolveThreads = 512;
static const size_t kSolveSmemMax = 227 * 1024;

namespace {

__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();  // red may still be read from a previous call
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < nw; ++w) s += red[w];
  return s;
}

// A <- H A H, H = I - tau v v^T (A symmetric, full storage)
__device__ void sym_reflect(double* A, int m, const double* v, double tau, double* w, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = warp; i < m; i += nw) {
    double s = 0.0;
    const double* row = A + (size_t)i * m;
    for (int j = lane; j < m; j += 32) s = fma(row[j], v[j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) w[i] = s;
  }
  __syncthreads();
  double part = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) part = fma(v[i], w[i], part);
  const double alpha = block_sum(part, red);
  const double c = 0.5 * tau * tau * alpha;
  for (int i = threadIdx.x; i < m; i += blockDim.x) w[i] = tau * w[i] - c * v[i];
  __syncthreads();
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int i = e / m, j = e - i * m;
    A[e] -= v[i] * w[j] + w[i] * v[j];
  }
  __syncthreads();
}

// x <- H x
__device__ void vec_reflect(double* x, int m, const double* v, double tau, double* red) {
  double part = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) part = fma(v[i], x[i], part);
  const double s = tau * block_sum(part, red);
  for (int i = threadIdx.x; i < m; i += blockDim.x) x[i] -= s * v[i];
  __syncthreads();
}

// Blocked algorithms on an m x m matrix that lives in shared memory (small systems) or in the
// L2-resident workspace (larger ones): the diagonal block D (kNB x kNB) and the current panel P
// (rows x kNB) are staged in shared memory, so the O(m^3) work reads its operands from shared memory
// and touches every trailing element once per block column instead of once per column.
static const int kNB = 32;
static const int kLD = kNB + 1;  // padded leading dimension of D and P: conflict-free column access

// Rows of the panel P that are staged in shared memory.  All of them up to m = 687 (every BASELINE model: m <= 601);
// a larger system keeps as many as fit next to its vectors and reads the rest of the panel from the (L2-resident)
// matrix itself: slower per element, same arithmetic in the same order.
__host__ __device__ inline int solve_panel_rows(int m) {  // solve_kernel: 4 m + 34 doubles of vectors, D, >= 32 columns of Y
  const long long room = 227ll * 1024 - 1024 - ((long long)4 * m + 34 + kNB * kLD + kNB * 32) * 8;
  const long long rows = room / (kLD * 8);
  return rows >= m ? m : (rows > 0 ? (int)rows : 0);
}
__host__ __device__ inline int chol_panel_rows(int m) {  // chol_check_kernel: D and the flag
  const long long rows = (227ll * 1024 - 1024 - ((long long)kNB * kLD + 2) * 8) / (kLD * 8);
  return rows >= m ? m : (int)rows;
}

// in-place lower Cholesky A = L L^T (upper part left untouched); sets *flag = 1 when A is not
// positive definite
__device__ void cholesky_lower(double* A, int m, double* D, double* P, int PR, int* flag) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  for (int k0 = 0; k0 < m; k0 += kNB) {
    const int kb = min(kNB, m - k0);
    for (int e = tid; e < kb * kb; e += nt) {
      const int i = e / kb, j = e - i * kb;
      D[i * kLD + j] = A[(size_t)(k0 + i) * m + k0 + j];
    }
    __syncthreads();
    if (warp == 0) {  // unblocked factorisation of the diagonal block: lane = row
      for (int k = 0; k < kb; ++k) {
        const double akk = D[k * kLD + k];
        if (!(akk > 0.0) || !isfinite(akk)) {
          if (lane == 0) *flag = 1;
          break;
        }
        const double d = sqrt(akk), inv = 1.0 / d;
        __syncwarp();
        if (lane >= k && lane < kb) D[lane * kLD + k] = (lane == k) ? d : D[lane * kLD + k] * inv;
        __syncwarp();
        if (lane > k && lane < kb) {
          const double lik = D[lane * kLD + k];
          for (int j = k + 1; j <= lane; ++j) D[lane * kLD + j] = fma(-lik, D[j * kLD + k], D[lane * kLD + j]);
        }
        __syncwarp();
      }
    }
    __syncthreads();
    if (*flag != 0) return;
    for (int e = tid; e < kb * kb; e += nt) {
      const int i = e / kb, j = e - i * kb;
      if (j <= i) A[(size_t)(k0 + i) * m + k0 + j] = D[i * kLD + j];
    }
    // panel: L21 = A21 L11^-T, one thread per row
    const int rows = m - k0 - kb;
    for (int ii = tid; ii < rows; ii += nt) {
      double* arow = A + (size_t)(k0 + kb + ii) * m + k0;
      double* prow = ii < PR ? P + (size_t)ii * kLD : arow;  // rows beyond the staged panel: in place
      for (int c = 0; c < kb; ++c) {
        double sacc = arow[c];
        for (int d = 0; d < c; ++d) sacc = fma(-prow[d], D[c * kLD + d], sacc);
        sacc /= D[c * kLD + c];
        prow[c] = sacc;
        arow[c] = sacc;
      }
    }
    __syncthreads();
    // trailing update of the lower triangle: A22 -= L21 L21^T
    for (int e = tid; e < rows * rows; e += nt) {
      const int ii = e / rows, jj = e - ii * rows;
      if (jj <= ii) {
        const double* pi = ii < PR ? P + (size_t)ii * kLD : A + (size_t)(k0 + kb + ii) * m + k0;
        const double* pj = jj < PR ? P + (size_t)jj * kLD : A + (size_t)(k0 + kb + jj) * m + k0;
        double sacc = 0.0;
        for (int c = 0; c < kb; ++c) sacc = fma(pi[c], pj[c], sacc);
        A[(size_t)(k0 + kb + ii) * m + k0 + kb + jj] -= sacc;
      }
    }
    __syncthreads();
  }
}

// x <- L^-1 x (forward) or L^-T x (backward), single right-hand side, blocked: the diagonal block
// is solved by one warp out of shared memory, the rest of x is updated by all threads
__device__ void tri_solve_vec(const double* L, int m, double* x, bool transpose, double* D) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = (m + kNB - 1) / kNB;
  for (int bi = 0; bi < nblk; ++bi) {
    const int k0 = (transpose ? (nblk - 1 - bi) : bi) * kNB;
    const int kb = min(kNB, m - k0);
    for (int e = tid; e < kb * kb; e += nt) {
      const int i = e / kb, j = e - i * kb;
      D[i * kLD + j] = L[(size_t)(k0 + i) * m + k0 + j];
    }
    __syncthreads();
    if (warp == 0) {
      double xi = (lane < kb) ? x[k0 + lane] : 0.0;
      if (!transpose) {
        for (int k = 0; k < kb; ++k) {
          const double xk = __shfl_sync(0xffffffffu, xi, k) / D[k * kLD + k];
          if (lane == k) xi = xk;
          else if (lane > k && lane < kb) xi = fma(-D[lane * kLD + k], xk, xi);
        }
      } else {
        for (int k = kb - 1; k >= 0; --k) {
          const double xk = __shfl_sync(0xffffffffu, xi, k) / D[k * kLD + k];
          if (lane == k) xi = xk;
          else if (lane < k) xi = fma(-D[k * kLD + lane], xk, xi);
        }
      }
      if (lane < kb) x[k0 + lane] = xi;
    }
    __syncthreads();
    if (!transpose) {
      for (int i = k0 + kb + tid; i < m; i += nt) {
        const double* Li = L + (size_t)i * m + k0;
        double sacc = x[i];
        for (int c = 0; c < kb; ++c) sacc = fma(-Li[c], x[k0 + c], sacc);
        x[i] = sacc;
      }
    } else {
      for (int i = tid; i < k0; i += nt) {
        double sacc = x[i];
        for (int c = 0; c < kb; ++c) sacc = fma(-L[(size_t)(k0 + c) * m + i], x[k0 + c], sacc);
        x[i] = sacc;
      }
    }
    __syncthreads();
  }
}

// B <- L^-1 B (transpose = false) or L^-T B (true), blocked: per block row the diagonal block D,
// the solved block Y (kNB x wc columns) and the panel of L that updates the remaining rows are in
// shared memory; the columns of B are processed in chunks of wc (what fits next to the panel)
__device__ void tri_solve_cols(const double* L, int m, double* B, bool transpose, double* D, double* P, int PR, double* Y,
                               int wc) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int nblk = (m + kNB - 1) / kNB;
  for (int j0 = 0; j0 < m; j0 += wc) {
    const int nc = min(wc, m - j0);
    for (int bi = 0; bi < nblk; ++bi) {
      const int k0 = (transpose ? (nblk - 1 - bi) : bi) * kNB;
      const int kb = min(kNB, m - k0);
      for (int e = tid; e < kb * kb; e += nt) {
        const int i = e / kb, j = e - i * kb;
        D[i * kLD + j] = L[(size_t)(k0 + i) * m + k0 + j];
      }
      // panel: the rows of B still to be updated, as P[row][c]
      const int prow0 = transpose ? 0 : k0 + kb, prows = transpose ? k0 : m - k0 - kb;
      const int srows = min(prows, PR);  // staged rows of the panel; the others are read from L
      for (int e = tid; e < srows * kb; e += nt) {
        if (!transpose) {
          const int ii = e / kb, c = e - ii * kb;
          P[(size_t)ii * kLD + c] = L[(size_t)(prow0 + ii) * m + k0 + c];
        } else {
          const int c = e / srows, ii = e - c * srows;  // consecutive threads read consecutive columns of L
          P[(size_t)ii * kLD + c] = L[(size_t)(k0 + c) * m + ii];
        }
      }
      __syncthreads();
      // block solve, one thread per column
      for (int j = tid; j < nc; j += nt) {
        double* bcol = B + (size_t)k0 * m + j0 + j;
        if (!transpose) {
          for (int c = 0; c < kb; ++c) {
            double sacc = bcol[(size_t)c * m];
            for (int d = 0; d < c; ++d) sacc = fma(-D[c * kLD + d], Y[(size_t)d * wc + j], sacc);
            sacc /= D[c * kLD + c];
            Y[(size_t)c * wc + j] = sacc;
            bcol[(size_t)c * m] = sacc;
          }
        } else {
          for (int c = kb - 1; c >= 0; --c) {
            double sacc = bcol[(size_t)c * m];
            for (int d = c + 1; d < kb; ++d) sacc = fma(-D[d * kLD + c], Y[(size_t)d * wc + j], sacc);
            sacc /= D[c * kLD + c];
            Y[(size_t)c * wc + j] = sacc;
            bcol[(size_t)c * m] = sacc;
          }
        }
      }
      __syncthreads();
      // update of the remaining rows: B[row][j] -= sum_c P[row][c] Y[c][j]
      for (int e = tid; e < prows * nc; e += nt) {
        const int ii = e / nc, j = e - ii * nc;
        const double* pr = P + (size_t)ii * kLD;
        size_t ps = 1;
        if (ii >= srows) {
          pr = transpose ? L + (size_t)k0 * m + ii : L + (size_t)(prow0 + ii) * m + k0;
          ps = transpose ? (size_t)m : 1;
        }
        double sacc = 0.0;
        for (int c = 0; c < kb; ++c) sacc = fma(pr[c * ps], Y[(size_t)c * wc + j], sacc);
        B[(size_t)(prow0 + ii) * m + j0 + j] -= sacc;
      }
      __syncthreads();
    }
  }
}

__device__ void transpose_inplace(double* B, int m) {
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int i = e / m, j = e - i * m;
    if (j > i) {
      const double a = B[e], b = B[(size_t)j * m + i];
      B[e] = b;
      B[(size_t)j * m + i] = a;
    }
  }
  __syncthreads();
}

}  // namespace

#define PGB_SOLVE_WANT_EDOF PGB_SOLVE_EDOF
#define PGB_SOLVE_WANT_COV PGB_SOLVE_COV
#define PGB_SOLVE_FACTORED 0x10000  // internal: the multi-CTA front (below) left L in A, [Q^T G Q]_0 in B, its status after B

__global__ void __launch_bounds__(kSolveThreads)
solve_kernel(int m, const double* __restrict__ G_all, const double* __restrict__ r_all, int share_gram,
             const double* __restrict__ EtE_all, const double* __restrict__ Enull_all,
             const double* __restrict__ hv,
             const double* __restrict__ tau, int p, const double* __restrict__ beta_prev_all,
             int share_prev, double* __restrict__ out_all, double* __restrict__ cov_all,
             double* __restrict__ ws_all, int flags, int use_smem, int wc) {
  extern __shared__ __align__(16) double sm[];
  const int sys = blockIdx.x;
  const size_t mm = (size_t)m * m;
  const double* G = G_all + (share_gram ? 0 : (size_t)sys * mm);
  const double* r = r_all + (share_gram ? 0 : (size_t)sys * m);
  const double* EtE = EtE_all + (size_t)sys * mm;
  const double* Enull = Enull_all ? Enull_all + (size_t)sys * mm : nullptr;
  const double* beta_prev = beta_prev_all ? beta_prev_all + (share_prev ? 0 : (size_t)sys * m) : nullptr;
  double* out = out_all + (size_t)sys * (m + PGB_SOLVE_OUT_EXTRA);
  // small vectors always live in shared memory
  double* v = sm;          // current reflector
  double* w = v + m;       // scratch
  double* x = w + m;       // right-hand side / solution
  double* col = x + m;     // Cholesky column
  double* red = col + m;   // 32 doubles
  int* flag = reinterpret_cast<int*>(red + 32);
  double* D = red + 34;                       // diagonal block of the blocked algorithms
  double* P = D + kNB * kLD;                  // panel, PR x kLD
  const int PR = solve_panel_rows(m);
  double* Y = P + (size_t)PR * kLD;           // solved block of the column solves, kNB x wc
  double* A = use_smem ? (Y + (size_t)kNB * wc) : (ws_all + (size_t)sys * 2 * mm);
  double* B = A + mm;
  const bool factored = (flags & PGB_SOLVE_FACTORED) != 0;
  if (threadIdx.x == 0) *flag = factored ? reinterpret_cast<const int*>(B + mm)[0] : 0;
  __syncthreads();

  // ---- A = Q^T E_rest Q + [Q^T E_null Q]_0 + [Q^T G Q]_0, where [.]_0 zeroes the first p rows and
  //      columns: both E_null and G annihilate the null vectors in exact arithmetic.  B ends up
  //      holding [Q^T G Q]_0 for the edof / cov step. ----
  int bad = 0;
  for (int e = threadIdx.x; e < (factored ? 0 : (int)mm); e += blockDim.x) {
    const double pe = EtE[e];
    if (!isfinite(pe)) bad = 1;
    A[e] = pe;
  }
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    x[i] = r[i];
    if (!isfinite(x[i])) bad = 1;
  }
  if (bad) *flag = 2;
  __syncthreads();
  const bool prepared = (flags & PGB_SOLVE_PREPARED) != 0;  // EtE already holds Q^T E_rest Q + [Q^T E_null Q]_0
  for (int q = 0; q < p && *flag == 0; ++q) {
    for (int i = threadIdx.x; i < m; i += blockDim.x) v[i] = hv[(size_t)q * m + i];
    __syncthreads();
    if (!prepared) sym_reflect(A, m, v, tau[q], w, red);
    vec_reflect(x, m, v, tau[q], red);
  }
  for (int pass = factored ? 2 : ((Enull && !prepared) ? 0 : 1); pass < 2; ++pass) {
    const double* src = (pass == 0) ? Enull : G;
    bad = 0;
    for (int e = threadIdx.x; e < (int)mm; e += blockDim.x) {
      const double g = src[e];
      if (!isfinite(g)) bad = 1;
      B[e] = g;
    }
    if (bad) *flag = 2;
    __syncthreads();
    if (*flag != 0) break;
    for (int q = 0; q < p; ++q) {
      for (int i = threadIdx.x; i < m; i += blockDim.x) v[i] = hv[(size_t)q * m + i];
      __syncthreads();
      sym_reflect(B, m, v, tau[q], w, red);
    }
    for (int e = threadIdx.x; e < (int)mm; e += blockDim.x) {
      const int i = e / m, j = e - i * m;
      double g = B[e];
      if (i < p || j < p) { g = 0.0; B[e] = 0.0; }
      A[e] += g;
    }
    __syncthreads();
  }
  if (*flag == 0) {
    for (int i = threadIdx.x; i < p; i += blockDim.x) x[i] = 0.0;
    __syncthreads();
    if (!factored) cholesky_lower(A, m, D, P, PR, flag);
  }
  __syncthreads();
  const int info = *flag;
  if (info != 0) {
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = threadIdx.x; i < m + PGB_SOLVE_OUT_EXTRA; i += blockDim.x) out[i] = nanv;
    __syncthreads();
    if (threadIdx.x == 0) out[m + 2] = (double)info;
    return;
  }
  // ---- beta = Q L^-T L^-1 x ----
  double ld_part = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) ld_part += log(A[(size_t)i * m + i]);
  const double logdet = 2.0 * block_sum(ld_part, red);
  tri_solve_vec(A, m, x, false, D);
  tri_solve_vec(A, m, x, true, D);
  for (int q = p - 1; q >= 0; --q) {
    for (int i = threadIdx.x; i < m; i += blockDim.x) v[i] = hv[(size_t)q * m + i];
    __syncthreads();
    vec_reflect(x, m, v, tau[q], red);
  }
  double n2 = 0.0, d2 = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const double b = x[i];
    out[i] = b;
    n2 = fma(b, b, n2);
    if (beta_prev) { const double d = beta_prev[i] - b; d2 = fma(d, d, d2); }
  }
  n2 = block_sum(n2, red);
  d2 = block_sum(d2, red);
  double edof = __longlong_as_double(0x7ff8000000000000LL);
  if (flags & (PGB_SOLVE_WANT_EDOF | PGB_SOLVE_WANT_COV)) {
    // Z = L^-1 Gt L^-T (symmetric); edof = tr(Z)
    tri_solve_cols(A, m, B, false, D, P, PR, Y, wc);
    transpose_inplace(B, m);
    tri_solve_cols(A, m, B, false, D, P, PR, Y, wc);
    double tr = 0.0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) tr += B[(size_t)i * m + i];
    edof = block_sum(tr, red);
    if ((flags & PGB_SOLVE_WANT_COV) && cov_all) {
      // cov = Q L^-T Z L^-1 Q^T
      tri_solve_cols(A, m, B, true, D, P, PR, Y, wc);
      transpose_inplace(B, m);
      tri_solve_cols(A, m, B, true, D, P, PR, Y, wc);
      for (int q = p - 1; q >= 0; --q) {
        for (int i = threadIdx.x; i < m; i += blockDim.x) v[i] = hv[(size_t)q * m + i];
        __syncthreads();
        sym_reflect(B, m, v, tau[q], w, red);
      }
      double* cov = cov_all + (size_t)sys * mm;
      for (int e = threadIdx.x; e < (int)mm; e += blockDim.x) cov[e] = B[e];
    }
  }
  if (threadIdx.x == 0) {
    out[m + 0] = edof;
    out[m + 1] = beta_prev ? sqrt(d2) / sqrt(n2) : __longlong_as_double(0x7ff8000000000000LL);
    out[m + 2] = 0.0;
    out[m + 3] = logdet;
    out[m + 4] = sqrt(n2);
    out[m + 5] = 0.0;
    out[m + 6] = 0.0;
    out[m + 7] = 0.0;
  }
}

// ----------------------------------------------------------------------------------------
// Multi-CTA front of a single large system (the PIRLS iteration of config 2 / 3: m = 251 / 601, two
// m x m matrices do not fit one SM's shared memory).  The single-CTA kernel above then works out of the
// L2 at one SM's latency; here every O(m^2 p) and O(m^3) sweep is a grid-wide kernel on the same stream:
//   B = G, A = A0                                   (front_init_kernel)
//   B <- H_q B H_q, q = 0 .. p-1                    (front_matvec_kernel + front_rank2_kernel per reflector)
//   A += [B]_0                                      (front_add_kernel)
//   A = L L^T, 32-wide block columns                (front_chol_panel_kernel + front_chol_trail_kernel per block)
// and the single-CTA kernel finishes (right-hand side, triangular solves, diff; edof / cov when asked).
// ----------------------------------------------------------------------------------------
static const int kFrontThreads = 256;

__global__ void front_init_kernel(int m, const double* __restrict__ G, const double* __restrict__ A0, double* __restrict__ A,
                                  double* __restrict__ B, int* __restrict__ status) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m * m) return;
  const double g = G[e], a = A0[e];
  if (!isfinite(g) || !isfinite(a)) *status = 2;
  B[e] = g;
  A[e] = a;
}

// w = B v, one warp per row
__global__ void front_matvec_kernel(int m, const double* __restrict__ B, const double* __restrict__ v, double* __restrict__ w) {
  const int lane = threadIdx.x & 31, i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= m) return;
  const double* row = B + (size_t)i * m;
  double s = 0.0;
  for (int j = lane; j < m; j += 32) s = fma(row[j], v[j], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) w[i] = s;
}

// B <- B - v w'^T - w' v^T with w' = tau w - (tau^2 (v^T w) / 2) v  (the update of sym_reflect)
__global__ void front_rank2_kernel(int m, double* __restrict__ B, const double* __restrict__ v, const double* __restrict__ w,
                                   const double* __restrict__ tau_q) {
  __shared__ double red[32];
  double part = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) part = fma(v[i], w[i], part);
  const double alpha = block_sum(part, red);
  const double tau = *tau_q;
  const double c = 0.5 * tau * tau * alpha;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < m * m; e += gridDim.x * blockDim.x) {
    const int i = e / m, j = e - i * m;
    const double vi = v[i], vj = v[j];
    const double wi = tau * w[i] - c * vi, wj = tau * w[j] - c * vj;
    B[e] -= vi * wj + wi * vj;
  }
}

// A += [B]_0: the first p rows and columns of B (the null directions) are exact zeros
__global__ void front_add_kernel(int m, int p, double* __restrict__ A, double* __restrict__ B) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m * m) return;
  const int i = e / m, j = e - i * m;
  double g = B[e];
  if (i < p || j < p) { g = 0.0; B[e] = 0.0; }
  A[e] += g;
}

// Cholesky of one 32 x 32 block by one warp, in registers: lane i holds row i; column k of L is sent
// around with shuffles.  Returns false when a pivot is not positive (or not finite).
__device__ __forceinline__ bool chol32_warp(double (&row)[kNB], int lane) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < kNB; ++k) {
    const double akk = __shfl_sync(0xffffffffu, row[k], k);
    if (!(akk > 0.0) || !isfinite(akk)) ok = false;
    const double d = sqrt(akk), inv = 1.0 / d;
    if (lane == k) row[k] = d;
    else if (lane > k) row[k] *= inv;
#pragma unroll
    for (int j = k + 1; j < kNB; ++j) {
      const double ljk = __shfl_sync(0xffffffffu, row[k], j);
      if (lane >= j) row[j] = fma(-row[k], ljk, row[j]);
    }
  }
  return ok;
}

// block column k0: every CTA factors the diagonal block (redundantly, one warp) and solves its slice of
// panel rows, L21 = A21 L11^-T; CTA 0 writes the diagonal block back
__global__ void __launch_bounds__(kFrontThreads)
front_chol_panel_kernel(int m, int k0, double* __restrict__ A, int* __restrict__ status) {
  __shared__ double D[kNB * kLD];
  __shared__ int bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (*status != 0) return;
  const int kb = min(kNB, m - k0);
  if (warp == 0) {
    // a short last block is padded with the identity
    double row[kNB];
#pragma unroll
    for (int j = 0; j < kNB; ++j)
      row[j] = (lane < kb && j < kb) ? ((j <= lane) ? A[(size_t)(k0 + lane) * m + k0 + j] : 0.0) : ((j == lane) ? 1.0 : 0.0);
    const bool ok = chol32_warp(row, lane);
    if (lane == 0) bad = ok ? 0 : 1;
#pragma unroll
    for (int j = 0; j < kNB; ++j) D[lane * kLD + j] = row[j];
  }
  __syncthreads();
  if (bad) {
    if (blockIdx.x == 0 && tid == 0) *status = 1;
    return;
  }
  if (blockIdx.x == 0)
    for (int e = tid; e < kb * kb; e += blockDim.x) {
      const int i = e / kb, j = e - i * kb;
      if (j <= i) A[(size_t)(k0 + i) * m + k0 + j] = D[i * kLD + j];
    }
  const int rows = m - k0 - kb;
  const int ii = blockIdx.x * blockDim.x + tid;
  if (ii < rows) {
    // row of L21: forward substitution, right-looking (the updates of one step are independent)
    double* arow = A + (size_t)(k0 + kb + ii) * m + k0;
    double pr[kNB];
#pragma unroll
    for (int c = 0; c < kNB; ++c) pr[c] = (c < kb) ? arow[c] : 0.0;
#pragma unroll
    for (int d = 0; d < kNB; ++d) {
      pr[d] /= D[d * kLD + d];
#pragma unroll
      for (int c = d + 1; c < kNB; ++c) pr[c] = fma(-pr[d], D[c * kLD + d], pr[c]);
    }
#pragma unroll
    for (int c = 0; c < kNB; ++c)
      if (c < kb) arow[c] = pr[c];
  }
}

// trailing update of the lower triangle, A22 -= L21 L21^T, one 32 x 32 tile per CTA
__global__ void __launch_bounds__(kFrontThreads)
front_chol_trail_kernel(int m, int k0, double* __restrict__ A, const int* __restrict__ status) {
  __shared__ double Pi[kNB * kLD], Pj[kNB * kLD];
  if (*status != 0) return;
  const int kb = min(kNB, m - k0), base = k0 + kb, rows = m - base;
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (tj > ti) return;
  const int i0 = ti * kNB, j0 = tj * kNB;
  for (int e = threadIdx.x; e < kNB * kNB; e += blockDim.x) {
    const int r = e / kNB, c = e - r * kNB;
    Pi[r * kLD + c] = (i0 + r < rows && c < kb) ? A[(size_t)(base + i0 + r) * m + k0 + c] : 0.0;
    Pj[r * kLD + c] = (j0 + r < rows && c < kb) ? A[(size_t)(base + j0 + r) * m + k0 + c] : 0.0;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kNB * kNB; e += blockDim.x) {
    const int r = e / kNB, c = e - r * kNB;
    const int gi = i0 + r, gj = j0 + c;
    if (gi < rows && gj <= gi) {
      double sacc = 0.0;
#pragma unroll
      for (int k = 0; k < kNB; ++k) sacc = fma(Pi[r * kLD + k], Pj[c * kLD + k], sacc);
      A[(size_t)(base + gi) * m + base + gj] -= sacc;
    }
  }
}

static int solve_front(int m, const double* G, const double* A0, const double* hv, const double* tau, int p, double* ws,
                       cudaStream_t st) {
  const size_t mm = (size_t)m * m;
  double* A = ws;
  double* B = A + mm;
  int* status = reinterpret_cast<int*>(B + mm);
  double* w = B + mm + 2;
  PGB_CUDA(cudaMemsetAsync(status, 0, 8, st));
  const int eg = (int)((mm + kFrontThreads - 1) / kFrontThreads);
  front_init_kernel<<<eg, kFrontThreads, 0, st>>>(m, G, A0, A, B, status);
  const int wpb = kFrontThreads / 32;
  for (int q = 0; q < p; ++q) {
    front_matvec_kernel<<<(m + wpb - 1) / wpb, kFrontThreads, 0, st>>>(m, B, hv + (size_t)q * m, w);
    front_rank2_kernel<<<std::min(eg, 296), kFrontThreads, 0, st>>>(m, B, hv + (size_t)q * m, w, tau + q);
  }
  front_add_kernel<<<eg, kFrontThreads, 0, st>>>(m, p, A, B);
  int launches = 2 + 2 * p + 1;
  for (int k0 = 0; k0 < m; k0 += kNB) {
    const int kb = std::min(kNB, m - k0), rows = m - k0 - kb;
    front_chol_panel_kernel<<<std::max(1, (rows + kFrontThreads - 1) / kFrontThreads), kFrontThreads, 0, st>>>(m, k0, A, status);
    ++launches;
    if (rows > 0) {
      const int nt = (rows + kNB - 1) / kNB;
      front_chol_trail_kernel<<<dim3(nt, nt), kFrontThreads, 0, st>>>(m, k0, A, status);
      ++launches;
    }
  }
  pgb_count_launch(launches);
  PGB_CUDA(cudaGetLastError());
  return PGB_OK;
}

// A0 = Q^T E_rest Q + [Q^T E_null Q]_0: the penalty side of the system in the null-space basis.
// It does not change during a PIRLS loop, so it is computed once per penalty (pgb_solve_prepare)
// and handed to pgb_solve with PGB_SOLVE_PREPARED.
__global__ void __launch_bounds__(kSolveThreads)
solve_prepare_kernel(int m, const double* __restrict__ EtE_all, const double* __restrict__ Enull_all,
                     const double* __restrict__ hv, const double* __restrict__ tau, int p, double* __restrict__ A0_all,
                     double* __restrict__ ws_all) {
  extern __shared__ __align__(16) double sm[];
  const int sys = blockIdx.x;
  const size_t mm = (size_t)m * m;
  double* v = sm;
  double* w = v + m;
  double* red = w + m;
  double* A = A0_all + (size_t)sys * mm;
  double* B = ws_all + (size_t)sys * mm;
  const double* EtE = EtE_all + (size_t)sys * mm;
  for (int e = threadIdx.x; e < (int)mm; e += blockDim.x) A[e] = EtE[e];
  if (Enull_all)
    for (int e = threadIdx.x; e < (int)mm; e += blockDim.x) B[e] = Enull_all[(size_t)sys * mm + e];
  __syncthreads();
  for (int q = 0; q < p; ++q) {
    for (int i = threadIdx.x; i < m; i += blockDim.x) v[i] = hv[(size_t)q * m + i];
    __syncthreads();
    sym_reflect(A, m, v, tau[q], w, red);
    if (Enull_all) sym_reflect(B, m, v, tau[q], w, red);
  }
  if (Enull_all) {
    for (int e = threadIdx.x; e < (int)mm; e += blockDim.x) {
      const int i = e / m, j = e - i * m;
      if (i >= p && j >= p) A[e] += B[e];
    }
  }
}

extern "C" int pgb_solve_prepare(int32_t m, int32_t nb, const double* EtE, const double* EtE_null, const double* hv,
                                 const double* tau, int32_t p, double* A0, void* ws, int64_t ws_bytes, void* stream) {
  if (m < 1 || nb < 1 || !EtE || !A0 || p < 0 || p > m || (p > 0 && (!hv || !tau)) ||
      (EtE_null && (!ws || ws_bytes < (int64_t)nb * m * m * 8))) {
    pgb_set_error("pgb_solve_prepare: bad arguments");
    return PGB_EINVAL;
  }
  solve_prepare_kernel<<<nb, kSolveThreads, ((size_t)2 * m + 34) * 8, (cudaStream_t)stream>>>(m, EtE, EtE_null, hv, tau, p, A0,
                                                                                              (double*)ws);
  pgb_count_launch(1);
  PGB_CUDA(cudaGetLastError());
  return PGB_OK;
}

// cached CUDA graphs of the multi-CTA front + finishing kernel, one per distinct argument set
struct SolveGraphKey {
  int32_t m, p, flags;
  const void *G, *r, *A0, *hv, *tau, *beta_prev, *out, *cov, *ws;
  bool operator==(const SolveGraphKey& o) const {
    return m == o.m && p == o.p && flags == o.flags && G == o.G && r == o.r && A0 == o.A0 && hv == o.hv && tau == o.tau &&
           beta_prev == o.beta_prev && out == o.out && cov == o.cov && ws == o.ws;
  }
};
struct SolveGraph {
  SolveGraphKey key;
  int dev, launches;
  cudaGraphExec_t exec;
};
static std::mutex g_graph_mu;
static std::vector<SolveGraph> g_graphs;
long long pgb_launch_count(void);

static size_t solve_vec_smem(int m) { return ((size_t)4 * m + 34 + kNB * kLD + (size_t)solve_panel_rows(m) * kLD) * 8; }
// columns of B solved at a time: as many as fit next to the vectors and the panel (at least 32)
static int solve_col_chunk(int m, bool with_matrices) {
  const size_t used = solve_vec_smem(m) + (with_matrices ? (size_t)2 * m * m * 8 : 0);
  if (used + (size_t)kNB * 32 * 8 > kSolveSmemMax) return 0;
  return (int)std::min<size_t>((size_t)m, (kSolveSmemMax - used) / ((size_t)kNB * 8));
}
static bool solve_fits_smem(int m) { return solve_col_chunk(m, true) >= std::min(m, 32); }

extern "C" int64_t pgb_solve_ws_bytes(int32_t m, int32_t nb) {
  if (solve_fits_smem(m)) return 256;
  return (int64_t)nb * 2 * m * m * 8 + (int64_t)(m + 4) * 8 + 256;  // A, B; status and one vector of the multi-CTA front
}

extern "C" int pgb_solve(int32_t m, int32_t nb, const double* G, const double* r, int32_t share_gram,
                         const double* EtE, const double* EtE_null, const double* hv, const double* tau,
                         int32_t p, const double* beta_prev, int32_t share_prev, int32_t flags, double* out,
                         double* cov_out, void* ws, int64_t ws_bytes, void* stream) {
  if (m < 1 || nb < 1 || !G || !r || !EtE || !out || p < 0 || p > m || (p > 0 && (!hv || !tau))) {
    pgb_set_error("pgb_solve: bad arguments");
    return PGB_EINVAL;
  }
  const bool in_smem = solve_fits_smem(m);
  if (!in_smem && (!ws || ws_bytes < pgb_solve_ws_bytes(m, nb))) {
    pgb_set_error("pgb_solve: workspace too small");
    return PGB_EINVAL;
  }
  const int wc = solve_col_chunk(m, in_smem);
  if (wc < 1) { pgb_set_error("pgb_solve: m=%d too large", m); return PGB_EUNSUPPORTED; }
  const size_t smem = solve_vec_smem(m) + (size_t)kNB * wc * 8 + (in_smem ? (size_t)2 * m * m * 8 : 0);
  PGB_CUDA(cudaFuncSetAttribute(solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSolveSmemMax));
  if (!in_smem && nb == 1 && (flags & PGB_SOLVE_PREPARED)) {
    // One large system: the front is ~40 short dependent kernels.  A PIRLS loop calls this with the same buffers
    // every iteration, so the whole sequence is captured once into a CUDA graph and replayed (the launch and
    // inter-kernel gaps were 15 % of a config-2 iteration, more in a strong-scaled run).
    const SolveGraphKey key{m, p, flags, G, r, EtE, hv, tau, beta_prev, out, cov_out, ws};
    int dev = 0;
    PGB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_graph_mu);
    for (SolveGraph& e : g_graphs)
      if (e.dev == dev && e.key == key) {
        PGB_CUDA(cudaGraphLaunch(e.exec, (cudaStream_t)stream));
        pgb_count_launch(e.launches);
        return PGB_OK;
      }
    cudaStream_t cs = nullptr;  // capture on a private stream: the caller's may be the legacy default stream
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool ok = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    int launches = 0;
    if (ok) {
      const long long before = pgb_launch_count();
      int rc = solve_front(m, G, EtE, hv, tau, p, (double*)ws, cs);
      solve_kernel<<<1, kSolveThreads, smem, cs>>>(m, G, r, share_gram, EtE, EtE_null, hv, tau, p, beta_prev, share_prev, out,
                                                   cov_out, (double*)ws, flags | PGB_SOLVE_FACTORED, 0, wc);
      launches = (int)(pgb_launch_count() - before) + 1;
      ok = cudaStreamEndCapture(cs, &graph) == cudaSuccess && rc == PGB_OK && graph != nullptr &&
           cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
    }
    if (graph) cudaGraphDestroy(graph);
    if (cs) cudaStreamDestroy(cs);
    if (ok) {
      if (g_graphs.size() >= 16) {  // a handful of engines per process; drop the oldest
        cudaGraphExecDestroy(g_graphs.front().exec);
        g_graphs.erase(g_graphs.begin());
      }
      g_graphs.push_back(SolveGraph{key, dev, launches, exec});
      PGB_CUDA(cudaGraphLaunch(exec, (cudaStream_t)stream));
      pgb_count_launch(1);
      return PGB_OK;
    }
    cudaGetLastError();  // capture not possible here: plain launches
    const int rc = solve_front(m, G, EtE, hv, tau, p, (double*)ws, (cudaStream_t)stream);
    if (rc != PGB_OK) return rc;
    flags |= PGB_SOLVE_FACTORED;
  }
  solve_kernel<<<nb, kSolveThreads, smem, (cudaStream_t)stream>>>(m, G, r, share_gram, EtE, EtE_null, hv, tau, p, beta_prev,
                                                                  share_prev, out, cov_out, (double*)ws, flags,
                                                                  in_smem ? 1 : 0, wc);
  pgb_count_launch(1);
  PGB_CUDA(cudaGetLastError());
  return PGB_OK;
}

// Cholesky test of S + P (+ C): GAM._cholesky (pygam.py:499-537) raises NotPositiveDefiniteError
__global__ void __launch_bounds__(kSolveThreads)
chol_check_kernel(int m, const double* __restrict__ Ain, int* __restrict__ info, double* __restrict__ ws,
                  int use_smem) {
  extern __shared__ __align__(16) double sm[];
  double* D = sm;
  double* P = D + kNB * kLD;
  const int PR = chol_panel_rows(m);
  int* flag = reinterpret_cast<int*>(P + (size_t)PR * kLD);
  const size_t sys = blockIdx.x;  // one CTA per matrix of a batch
  double* A = use_smem ? (P + (size_t)PR * kLD + 2) : ws + sys * (size_t)m * m;
  Ain += sys * (size_t)m * m;
  if (threadIdx.x == 0) *flag = 0;
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) A[e] = Ain[e];
  __syncthreads();
  cholesky_lower(A, m, D, P, PR, flag);
  __syncthreads();
  if (threadIdx.x == 0) info[sys] = *flag;
}


// ----------------------------------------------------------------------------------------
// Rank-limited solve: fewer kept rows than coefficients.
// The reference crops its SVD of [R; E] to the leading min(m, n, rows kept by the mask) singular triplets
// (pygam.py:789-799, exercised by pygam/tests/test_GAM_methods.py:93-107): with A = G + E^T E = V D^2 V^T that is
//   beta = V_k D_k^-2 V_k^T r,   edof = tr(D_k^-1 V_k^T G V_k D_k^-1),   cov = V_k D_k^-2 (V_k^T G V_k) D_k^-2 V_k^T
// over the k largest eigenvalues.  One CTA: two-sided cyclic Jacobi on A in global memory (round-robin pairs: the
// m / 2 rotations of a step touch disjoint rows and columns and are applied together), then the three formulas.
// m is small whenever this is needed (there are fewer rows than coefficients).
// ----------------------------------------------------------------------------------------
static const int kEigThreads = 512;
__global__ void __launch_bounds__(kEigThreads)
solve_truncated_kernel(int m, int rank, const double* __restrict__ G, const double* __restrict__ r,
                       const double* __restrict__ EtE, const double* __restrict__ EtE2, const double* __restrict__ beta_prev,
                       double* __restrict__ out, double* __restrict__ cov_out, double* __restrict__ ws, int want_cov) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, T = blockDim.x;
  const int me = (m + 1) & ~1;        // players of the round-robin (one dummy when m is odd)
  const int np2 = me / 2;
  double* cs = sm;                    // [np2][2] rotation of every pair of the step
  int* pq = reinterpret_cast<int*>(cs + 2 * np2);   // [np2][2]
  double* lam = reinterpret_cast<double*>(pq + 2 * np2 + (2 * np2 & 1 ? 1 : 0));  // [m]
  double* vr = lam + m;               // [m] V^T r / lambda
  int* rk = reinterpret_cast<int*>(vr + m);  // [m] rank of eigenvalue i in descending order
  double* red = reinterpret_cast<double*>(rk + m + (m & 1));  // [64]
  double* A = ws;                     // m x m
  double* V = A + (size_t)m * m;      // m x m, columns = eigenvectors
  double* M = V + (size_t)m * m;      // m x m scratch
  __shared__ int s_flag;
  int bad = 0;
  for (int e = tid; e < m * m; e += T) {
    const int i = e / m, j = e - i * m;
    const double a = G[e] + EtE[e] + (EtE2 ? EtE2[e] : 0.0);
    if (!isfinite(a)) bad = 1;
    A[e] = a;
    V[e] = (i == j) ? 1.0 : 0.0;
  }
  if (tid == 0) s_flag = 0;
  __syncthreads();
  if (bad) s_flag = 2;
  __syncthreads();
  // symmetrise (G is symmetric by construction; the penalty is assembled on the host)
  for (int e = tid; e < m * m; e += T) {
    const int i = e / m, j = e - i * m;
    if (i < j) { const double a = 0.5 * (A[e] + A[j * m + i]); A[e] = a; A[j * m + i] = a; }
  }
  __syncthreads();
  for (int sweep = 0; sweep < 40 && s_flag == 0; ++sweep) {
    // convergence: off-diagonal mass against the diagonal
    double off = 0.0, dg = 0.0;
    for (int e = tid; e < m * m; e += T) {
      const int i = e / m, j = e - i * m;
      const double a = A[e];
      if (i == j) dg += a * a; else off += a * a;
    }
    for (int o = 16; o > 0; o >>= 1) { off += __shfl_down_sync(0xffffffffu, off, o); dg += __shfl_down_sync(0xffffffffu, dg, o); }
    if ((tid & 31) == 0) { red[tid >> 5] = off; red[32 + (tid >> 5)] = dg; }
    __syncthreads();
    if (tid == 0) {
      double so = 0.0, sd = 0.0;
      for (int w = 0; w < T / 32; ++w) { so += red[w]; sd += red[32 + w]; }
      if (so <= 1e-30 * sd) s_flag = -1;  // converged
    }
    __syncthreads();
    if (s_flag != 0) break;
    for (int step = 0; step < me - 1; ++step) {
      // pairs of this step: player me - 1 stays, the others rotate
      for (int k = tid; k < np2; k += T) {
        int p = (k == 0) ? me - 1 : (step + k) % (me - 1);
        int q = (k == 0) ? step : (step - k + (me - 1)) % (me - 1);
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, sn = 0.0;
        if (q < m) {
          const double apq = A[p * m + q];
          if (apq != 0.0) {
            const double tau = (A[q * m + q] - A[p * m + p]) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            sn = t * c;
          }
        } else {
          p = -1;
        }
        cs[2 * k] = c; cs[2 * k + 1] = sn;
        pq[2 * k] = p; pq[2 * k + 1] = q;
      }
      __syncthreads();
      // columns: A <- A J, V <- V J
      for (int e = tid; e < np2 * m; e += T) {
        const int k = e / m, i = e - k * m;
        const int p = pq[2 * k], q = pq[2 * k + 1];
        if (p < 0) continue;
        const double c = cs[2 * k], sn = cs[2 * k + 1];
        const double aip = A[i * m + p], aiq = A[i * m + q];
        A[i * m + p] = c * aip - sn * aiq;
        A[i * m + q] = sn * aip + c * aiq;
        const double vip = V[i * m + p], viq = V[i * m + q];
        V[i * m + p] = c * vip - sn * viq;
        V[i * m + q] = sn * vip + c * viq;
      }
      __syncthreads();
      // rows: A <- J^T A
      for (int e = tid; e < np2 * m; e += T) {
        const int k = e / m, j = e - k * m;
        const int p = pq[2 * k], q = pq[2 * k + 1];
        if (p < 0) continue;
        const double c = cs[2 * k], sn = cs[2 * k + 1];
        const double apj = A[p * m + j], aqj = A[q * m + j];
        A[p * m + j] = c * apj - sn * aqj;
        A[q * m + j] = sn * apj + c * aqj;
      }
      __syncthreads();
    }
  }
  const int info = (s_flag == 2) ? 2 : 0;
  // eigenvalues, their order, V^T r / lambda
  for (int i = tid; i < m; i += T) lam[i] = A[i * m + i];
  __syncthreads();
  for (int i = tid; i < m; i += T) {
    int rnk = 0;
    for (int j = 0; j < m; ++j) rnk += (lam[j] > lam[i]) || (lam[j] == lam[i] && j < i);
    rk[i] = rnk;
    double d = 0.0;
    for (int j = 0; j < m; ++j) d += V[j * m + i] * r[j];
    vr[i] = (rnk < rank && lam[i] > 0.0) ? d / lam[i] : 0.0;
  }
  __syncthreads();
  // beta = V_k (V_k^T r / lambda_k)
  for (int i = tid; i < m; i += T) {
    double b = 0.0;
    for (int j = 0; j < m; ++j) b += V[i * m + j] * vr[j];
    out[i] = b;
  }
  // M = G V (columns of kept eigenvectors only), then W = V^T M scaled by 1 / (lambda_i lambda_j) into A
  for (int e = tid; e < m * m; e += T) {
    const int i = e / m, j = e - i * m;
    double a = 0.0;
    if (rk[j] < rank)
      for (int l = 0; l < m; ++l) a += G[i * m + l] * V[l * m + j];
    M[e] = a;
  }
  __syncthreads();
  for (int e = tid; e < m * m; e += T) {
    const int i = e / m, j = e - i * m;
    double a = 0.0;
    if (rk[i] < rank && rk[j] < rank && lam[i] > 0.0 && lam[j] > 0.0) {
      for (int l = 0; l < m; ++l) a += V[l * m + i] * M[l * m + j];
      a /= lam[i] * lam[j];
    }
    A[e] = a;  // Lambda^-1 V^T G V Lambda^-1 on the kept block
  }
  __syncthreads();
  if (tid == 0) {
    double edof = 0.0, nb2 = 0.0, nd2 = 0.0;
    for (int i = 0; i < m; ++i) {
      if (rk[i] < rank) edof += A[i * m + i] * lam[i];  // v^T G v / lambda
      nb2 += out[i] * out[i];
      if (beta_prev) { const double d = beta_prev[i] - out[i]; nd2 += d * d; }
    }
    out[m] = edof;
    out[m + 1] = beta_prev ? sqrt(nd2) / sqrt(nb2) : 0.0;
    out[m + 2] = (double)info;
    out[m + 3] = 0.0;
    out[m + 4] = sqrt(nb2);
  }
  if (want_cov) {
    // cov = V (Lambda^-1 V^T G V Lambda^-1) V^T: M = V A, cov = M V^T
    for (int e = tid; e < m * m; e += T) {
      const int i = e / m, j = e - i * m;
      double a = 0.0;
      for (int l = 0; l < m; ++l) a += V[i * m + l] * A[l * m + j];
      M[e] = a;
    }
    __syncthreads();
    for (int e = tid; e < m * m; e += T) {
      const int i = e / m, j = e - i * m;
      double a = 0.0;
      for (int l = 0; l < m; ++l) a += M[i * m + l] * V[j * m + l];
      cov_out[e] = a;
    }
  }
}

extern "C" int64_t pgb_solve_truncated_ws_bytes(int32_t m) { return (int64_t)3 * m * m * 8 + 256; }

extern "C" int pgb_solve_truncated(int32_t m, int32_t rank, const double* G, const double* r, const double* EtE,
                                   const double* EtE_null, const double* beta_prev, int32_t flags, double* out,
                                   double* cov_out, void* ws, int64_t ws_bytes, void* stream) {
  if (m < 1 || rank < 1 || rank > m || !G || !r || !EtE || !out || !ws || ws_bytes < pgb_solve_truncated_ws_bytes(m) ||
      ((flags & PGB_SOLVE_COV) && !cov_out)) {
    pgb_set_error("pgb_solve_truncated: bad arguments");
    return PGB_EINVAL;
  }
  const size_t smem = ((size_t)(m + 2) * 2 + 2 * m + 64 + 8) * 8 + (size_t)(3 * m + 8) * 4;
  solve_truncated_kernel<<<1, kEigThreads, smem, (cudaStream_t)stream>>>(m, rank, G, r, EtE, EtE_null, beta_prev, out, cov_out,
                                                                        (double*)ws, (flags & PGB_SOLVE_COV) ? 1 : 0);
  pgb_count_launch(1);
  PGB_CUDA(cudaGetLastError());
  return PGB_OK;
}

extern "C" int64_t pgb_chol_ws_bytes(int32_t m) { return (int64_t)m * m * 8 + 256; }

extern "C" int pgb_chol_check_batch(int32_t m, int32_t nb, const double* A, int32_t* info_out, void* ws, int64_t ws_bytes,
                                    void* stream) {
  if (m < 1 || nb < 1 || !A || !info_out) { pgb_set_error("pgb_chol_check: bad arguments"); return PGB_EINVAL; }
  const size_t vec = ((size_t)kNB * kLD + (size_t)chol_panel_rows(m) * kLD + 2) * 8;
  const bool in_smem = vec + (size_t)m * m * 8 <= kSolveSmemMax;
  if (!in_smem && (!ws || ws_bytes < (int64_t)nb * m * m * 8)) {
    pgb_set_error("pgb_chol_check: workspace too small");
    return PGB_EINVAL;
  }
  PGB_CUDA(cudaFuncSetAttribute(chol_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSolveSmemMax));
  chol_check_kernel<<<nb, kSolveThreads, vec + (in_smem ? (size_t)m * m * 8 : 0), (cudaStream_t)stream>>>(
      m, A, info_out, (double*)ws, in_smem ? 1 : 0);
  pgb_count_launch(1);
  PGB_CUDA(cudaGetLastError());
  return PGB_OK;
}

extern "C" int pgb_chol_check(int32_t m, const double* A, int32_t* info_out, void* ws, int64_t ws_bytes,
                              void* stream) {
  if (m < 1 || !A || !info_out) { pgb_set_error("pgb_chol_check: bad arguments"); return PGB_EINVAL; }
  const size_t vec = ((size_t)kNB * kLD + (size_t)chol_panel_rows(m) * kLD + 2) * 8;
  const bool in_smem = vec + (size_t)m * m * 8 <= kSolveSmemMax;
  if (!in_smem && (!ws || ws_bytes < pgb_chol_ws_bytes(m))) {
    pgb_set_error("pgb_chol_check: workspace too small");
    return PGB_EINVAL;
  }
  PGB_CUDA(cudaFuncSetAttribute(chol_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSolveSmemMax));
  chol_check_kernel<<<1, kSolveThreads, vec + (in_smem ? (size_t)m * m * 8 : 0), (cudaStream_t)stream>>>(
      m, A, info_out, (double*)ws, in_smem ? 1 : 0);
  pgb_count_launch(1);
  PGB_CUDA(cudaGetLastError());
  return PGB_OK;
}
