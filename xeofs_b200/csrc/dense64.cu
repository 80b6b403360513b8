// Dense fp64 kernels for the k-column algebra beyond one 128-column block and for the small matrices of the cross
// models' PCA stage: a general matrix product, the cross-block Gram matrix of a tall k-column matrix, and a
// multi-CTA one-sided Jacobi eigen-solver.  They stand in for what the reference gets from LAPACK inside
// sklearn.utils.extmath.randomized_svd (scipy.linalg.lu / qr / svd of the l-column iterates; call site
// xeofs/linalg/decomposer.py:141-146) and numpy (preprocessing/pca.py:94-131 via linalg/_numpy/_svd.py:141-202;
// the Whitener's and the rotators' m x m products) when l = n_modes + 10 exceeds 128.
#include <math.h>

#include "common.cuh"

namespace xb {

// ------------------------------------------------------------------------------------------------ C = alpha op(A) op(B) + beta C
// Row-major fp64.  64 x 64 tile of C per block, 256 threads x (4 x 4), K in steps of 16 through shared memory.
constexpr int DG_T = 64, DG_K = 16;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
dgemm_kernel(int64_t m, int64_t n, int64_t k, double alpha, const double* __restrict__ A, int64_t lda,
             const double* __restrict__ B, int64_t ldb, double beta, double* __restrict__ C, int64_t ldc) {
  __shared__ double As[DG_K][DG_T + 1];  // [kk][i]
  __shared__ double Bs[DG_K][DG_T + 1];  // [kk][j]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * DG_T, j0 = (int64_t)blockIdx.x * DG_T;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int64_t k0 = 0; k0 < k; k0 += DG_K) {
    // op(A)[i, kk]: A[i*lda + kk] or (TA) A[kk*lda + i];   op(B)[kk, j]: B[kk*ldb + j] or (TB) B[j*ldb + kk]
    for (int idx = tid; idx < DG_T * DG_K; idx += 256) {
      int i, kk;
      if (TA) { i = idx % DG_T; kk = idx / DG_T; } else { kk = idx % DG_K; i = idx / DG_K; }
      const int64_t gi = i0 + i, gk = k0 + kk;
      As[kk][i] = (gi < m && gk < k) ? (TA ? A[gk * lda + gi] : A[gi * lda + gk]) : 0.0;
      int j, kb;
      if (TB) { kb = idx % DG_K; j = idx / DG_K; } else { j = idx % DG_T; kb = idx / DG_T; }
      const int64_t gj = j0 + j, gkb = k0 + kb;
      Bs[kb][j] = (gj < n && gkb < k) ? (TB ? B[gj * ldb + gkb] : B[gkb * ldb + gj]) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < DG_K; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) { a[x] = As[kk][ty + 16 * x]; b[x] = Bs[kk][tx + 16 * x]; }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int64_t gi = i0 + ty + 16 * x, gj = j0 + tx + 16 * y;
      if (gi < m && gj < n) {
        const double old = beta != 0.0 ? beta * C[gi * ldc + gj] : 0.0;
        C[gi * ldc + gj] = alpha * acc[x][y] + old;
      }
    }
}

// ------------------------------------------------------------------------------------------------ cross-block Gram
// G[ia + i, ib + j] += sum_n M(n, ia + i) M(n, ib + j) for one pair of column blocks (wa, wb <= 128 columns), fp64
// accumulation of fp32 data; persistent blocks over chunks of n like gram_kernel, every thread an 8 x 8 register tile.
constexpr int G2_CHUNK = 16;

template <int SIDE>
__global__ void __launch_bounds__(256)
gram2_kernel(const float* __restrict__ M, int64_t n, int64_t ld, int ia, int wa, int ib, int wb, double* __restrict__ G,
             int64_t ldg, int mirror) {
  __shared__ double sa[G2_CHUNK][128 + 2], sb[G2_CHUNK][128 + 2];
  const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;  // tile (ti, tj): rows ti*8.., cols tj*8..
  double acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;
  const int64_t n_chunks = (n + G2_CHUNK - 1) / G2_CHUNK;
  for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const int64_t n0 = ch * G2_CHUNK;
    __syncthreads();
    for (int idx = tid; idx < G2_CHUNK * 128; idx += 256) {
      int r, j;
      if (SIDE == 1) { r = idx % G2_CHUNK; j = idx / G2_CHUNK; } else { j = idx % 128; r = idx / 128; }
      const bool in = n0 + r < n;
      const int64_t ea = SIDE == 1 ? (int64_t)(ia + j) * ld + n0 + r : (n0 + r) * ld + ia + j;
      const int64_t eb = SIDE == 1 ? (int64_t)(ib + j) * ld + n0 + r : (n0 + r) * ld + ib + j;
      sa[r][j] = (in && j < wa) ? (double)M[ea] : 0.0;
      sb[r][j] = (in && j < wb) ? (double)M[eb] : 0.0;
    }
    __syncthreads();
#pragma unroll 2
    for (int r = 0; r < G2_CHUNK; ++r) {
      double a[8], b[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) { a[x] = sa[r][ti * 8 + x]; b[x] = sb[r][tj * 8 + x]; }
#pragma unroll
      for (int x = 0; x < 8; ++x)
#pragma unroll
        for (int y = 0; y < 8; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
    }
  }
#pragma unroll
  for (int x = 0; x < 8; ++x)
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      const int i = ti * 8 + x, j = tj * 8 + y;
      if (i < wa && j < wb) {
        atomicAdd(&G[(int64_t)(ia + i) * ldg + ib + j], acc[x][y]);
        if (mirror) atomicAdd(&G[(int64_t)(ib + j) * ldg + ia + i], acc[x][y]);
      }
    }
}

// ------------------------------------------------------------------------------------------------ wide symmetric eigen-solver
// One-sided (Hestenes) Jacobi on the rows of a symmetric positive semi-definite matrix W = G (n x n): plane rotations
// J make the rows of J G mutually orthogonal; then the row norms are the eigenvalues and the accumulated rotations
// (rows of V) the eigenvectors.  One warp per row pair, n/2 disjoint pairs per round (round-robin seating), one
// launch per round.  `state`: [0] = largest |cos| between two rows met in the current sweep, [1] = converged flag.
__device__ __forceinline__ int jw_seat(int q, int r, int ne) {
  if (q == 0) return 0;
  int x = q - 1 + r;
  if (x >= ne - 1) x -= ne - 1;
  return 1 + x;
}

__global__ void __launch_bounds__(256)
jacobi_round_kernel(double* __restrict__ W, double* __restrict__ V, int n, int ne, int round, double* __restrict__ state) {
  if (state[1] != 0.0) return;  // converged in an earlier sweep
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= ne / 2) return;
  int p = jw_seat(warp, round, ne), r = jw_seat(ne - 1 - warp, round, ne);
  if (p > r) { const int t = p; p = r; r = t; }
  if (r >= n) return;  // the bye of an odd n
  double* wp = W + (int64_t)p * n;
  double* wr = W + (int64_t)r * n;
  double a = 0.0, b = 0.0, c = 0.0;
  for (int j = lane; j < n; j += 32) {
    const double x = wp[j], y = wr[j];
    a = fma(x, x, a); b = fma(y, y, b); c = fma(x, y, c);
  }
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  const double denom = sqrt(a * b);
  if (!(denom > 0.0) || !(fabs(c) > 1e-300)) return;
  const double cosang = fabs(c) / denom;
  if (lane == 0 && cosang > state[0]) atomicMax((unsigned long long*)&state[0], (unsigned long long)__double_as_longlong(cosang));
  if (cosang < 1e-15) return;
  // rotation that makes the two rows orthogonal (the smaller of the two angles)
  const double zeta = (b - a) / (2.0 * c);
  const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double cs = rsqrt(1.0 + t * t), sn = cs * t;
  for (int j = lane; j < n; j += 32) {
    const double x = wp[j], y = wr[j];
    wp[j] = cs * x - sn * y;
    wr[j] = sn * x + cs * y;
  }
  double* vp = V + (int64_t)p * n;
  double* vr = V + (int64_t)r * n;
  for (int j = lane; j < n; j += 32) {
    const double x = vp[j], y = vr[j];
    vp[j] = cs * x - sn * y;
    vr[j] = sn * x + cs * y;
  }
}

__global__ void jacobi_init_kernel(const double* __restrict__ G, double* __restrict__ W, double* __restrict__ V, int n,
                                   double* __restrict__ state) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0) { state[0] = 0.0; state[1] = 0.0; }
  if (idx >= (int64_t)n * n) return;
  const int i = (int)(idx / n), j = (int)(idx % n);
  W[idx] = 0.5 * (G[idx] + G[(int64_t)j * n + i]);
  V[idx] = i == j ? 1.0 : 0.0;
}
// end of a sweep: converged when no pair of rows was further from orthogonal than the threshold
__global__ void jacobi_sweep_end_kernel(double* __restrict__ state, int32_t* __restrict__ info) {
  if (state[1] == 0.0) {
    info[0] += 1;
    if (state[0] < 5e-14) state[1] = 1.0;
  }
  state[0] = 0.0;
}
// eigenvalue i = norm of row i of W, ranked descending; eigenvector i = row i of V, written as column rank(i)
__global__ void __launch_bounds__(256)
jacobi_finish_kernel(const double* __restrict__ W, const double* __restrict__ V, int n, double* __restrict__ norms,
                     double* __restrict__ evals, double* __restrict__ evecs, int phase) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  if (phase == 0) {
    double a = 0.0;
    for (int j = lane; j < n; j += 32) { const double x = W[(int64_t)warp * n + j]; a = fma(x, x, a); }
    a = warp_sum(a);
    if (lane == 0) norms[warp] = sqrt(a);
    return;
  }
  const double vi = norms[warp];
  int rank = 0;
  for (int j = lane; j < n; j += 32) {
    const double vj = norms[j];
    rank += (vj > vi) || (vj == vi && j < warp);
  }
  rank = warp_sum(rank);
  if (lane == 0) evals[rank] = vi;
  for (int j = lane; j < n; j += 32) evecs[(int64_t)j * n + rank] = V[(int64_t)warp * n + j];
}

}  // namespace xb

using namespace xb;

extern "C" int xeofs_b200_dgemm(int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                                int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc,
                                void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(A && B && C && m > 0 && n > 0 && k > 0, "dgemm: bad arguments");
  XB_CHECK_ARG(ceil_div(m, DG_T) <= 65535, "dgemm: m too large");
  dim3 grid((unsigned)ceil_div(n, DG_T), (unsigned)ceil_div(m, DG_T));
  if (trans_a && trans_b) dgemm_kernel<true, true><<<grid, 256, 0, stream>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (trans_a) dgemm_kernel<true, false><<<grid, 256, 0, stream>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (trans_b) dgemm_kernel<false, true><<<grid, 256, 0, stream>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  else dgemm_kernel<false, false><<<grid, 256, 0, stream>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

extern "C" int xeofs_b200_gram_wide(const float* M, int64_t n, int64_t l, int64_t ld, int side, double* G, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(M && G && n > 0 && l > 0 && l <= 4096, "gram_wide: bad arguments (l=%lld must be in 1..4096)", (long long)l);
  XB_CHECK_ARG(side == 0 || side == 1, "gram_wide: side must be 0 (time-side) or 1 (space-side)");
  XB_CUDA(cudaMemsetAsync(G, 0, (size_t)l * l * sizeof(double), stream));
  const int blocks = (int)imin(ceil_div(n, G2_CHUNK), 2 * (int64_t)num_sms());
  for (int ia = 0; ia < l; ia += 128)
    for (int ib = ia; ib < l; ib += 128) {
      const int wa = (int)imin(128, l - ia), wb = (int)imin(128, l - ib);
      if (side == 0) gram2_kernel<0><<<blocks, 256, 0, stream>>>(M, n, ld, ia, wa, ib, wb, G, l, ib > ia);
      else gram2_kernel<1><<<blocks, 256, 0, stream>>>(M, n, ld, ia, wa, ib, wb, G, l, ib > ia);
      XB_LAUNCH_CHECK();
    }
  return XEOFS_OK;
}

extern "C" int64_t xeofs_b200_sym_eig_wide_workspace_bytes(int64_t n) { return (2 * n * n + n + 8) * (int64_t)sizeof(double); }

extern "C" int xeofs_b200_sym_eig_wide(const double* G, int64_t n, double* evals, double* evecs, void* workspace,
                                       int64_t workspace_bytes, int32_t* info, int max_sweeps, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(G && evals && evecs && workspace && info && n > 1 && n <= 4096, "sym_eig_wide: bad arguments");
  XB_CHECK_ARG(workspace_bytes >= xeofs_b200_sym_eig_wide_workspace_bytes(n), "sym_eig_wide: workspace too small");
  double* W = (double*)workspace;
  double* V = W + n * n;
  double* norms = V + n * n;
  double* state = norms + n;
  const int ne = ((int)n + 1) & ~1;
  XB_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), stream));
  jacobi_init_kernel<<<(unsigned)ceil_div(n * n, 256), 256, 0, stream>>>(G, W, V, (int)n, state);
  XB_LAUNCH_CHECK();
  const unsigned blocks = (unsigned)ceil_div(ne / 2, 8);
  if (max_sweeps <= 0) max_sweeps = 16;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    for (int round = 0; round < ne - 1; ++round) {
      jacobi_round_kernel<<<blocks, 256, 0, stream>>>(W, V, (int)n, ne, round, state);
    }
    XB_LAUNCH_CHECK();
    jacobi_sweep_end_kernel<<<1, 1, 0, stream>>>(state, info);
  }
  const unsigned fb = (unsigned)ceil_div(n, 8);
  jacobi_finish_kernel<<<fb, 256, 0, stream>>>(W, V, (int)n, norms, evals, evecs, 0);
  jacobi_finish_kernel<<<fb, 256, 0, stream>>>(W, V, (int)n, norms, evals, evecs, 1);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

// ------------------------------------------------------------------------------------------------ varimax: the m x m step
// What follows the sweep over the loadings in one iteration of linalg/_numpy/_rotation.py:166-177, on the device and
// without a host round trip:  G = G3 - alpha (XtX R) diag(W);  U, svals, V^T = svd(G);  R <- U V^T;  delta = sum(svals),
// through the eigen-decomposition of G^T G taken in the eigenbasis of the previous iteration (where the matrix is
// nearly diagonal already: the Jacobi sweeps of sym_eig end after two or three).
namespace xb {
__global__ void vu_form_G_kernel(const double* __restrict__ G3, const double* __restrict__ T1, const double* __restrict__ W,
                                 double alpha, int m, double* __restrict__ G) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < m * m) G[idx] = G3[idx] - alpha * T1[idx] * W[idx % m];
}
// P = V diag(1 / sqrt(ev)),  delta = sum sqrt(ev)   (one block)
__global__ void __launch_bounds__(256)
vu_scale_kernel(const double* __restrict__ V, const double* __restrict__ ev, int m, double* __restrict__ P,
                double* __restrict__ dsum) {
  __shared__ double inv[128];
  __shared__ double part[8];
  const int tid = threadIdx.x;
  double s = 0.0;
  if (tid < m) {
    const double sv = sqrt(fmax(ev[tid], 0.0));
    inv[tid] = sv > 0.0 ? 1.0 / sv : 0.0;
    s = sv;
  }
  s = warp_sum(s);
  if ((tid & 31) == 0) part[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += part[i];
    *dsum = t;
  }
  for (int idx = tid; idx < m * m; idx += 256) P[idx] = V[idx] * inv[idx % m];
}
template <bool TA, bool TB>
static void dgemm_sq(int m, const double* A, const double* B, double* C, cudaStream_t stream) {
  dim3 grid((unsigned)ceil_div(m, DG_T), (unsigned)ceil_div(m, DG_T));
  dgemm_kernel<TA, TB><<<grid, 256, 0, stream>>>(m, m, m, 1.0, A, m, B, m, 0.0, C, m);
}
}  // namespace xb

namespace xb {
int sym_eig_launch(const double* G, int64_t l, double* evals, double* evecs, double* work, int32_t* info, double tol2,
                   cudaStream_t stream);
}

extern "C" int64_t xeofs_b200_varimax_update_workspace_bytes(int64_t m) {
  return (9 * m * m + m + (m + 2) * (m + 2) + 16) * (int64_t)sizeof(double);
}

extern "C" int xeofs_b200_varimax_update(const double* G3, const double* W, const double* XtX, double alpha, int64_t m,
                                         double* R, double* basis, double* dsum, double eig_tol, void* workspace,
                                         int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(G3 && W && XtX && R && basis && dsum && workspace && m >= 2 && m <= 128,
               "varimax_update: bad arguments (m=%lld must be in 2..128)", (long long)m);
  XB_CHECK_ARG(workspace_bytes >= xeofs_b200_varimax_update_workspace_bytes(m), "varimax_update: workspace too small");
  const int mi = (int)m;
  const int64_t mm = m * m;
  double* T1 = (double*)workspace;
  double* G = T1 + mm;
  double* M = G + mm;
  double* T2 = M + mm;
  double* Mp = T2 + mm;
  double* Vp = Mp + mm;
  double* V = Vp + mm;
  double* P = V + mm;
  double* T3 = P + mm;
  double* ev = T3 + mm;
  double* ework = ev + m;
  int32_t* info = (int32_t*)(ework + (m + 2) * (m + 2));
  const unsigned eb = (unsigned)ceil_div(mm, 256);
  dgemm_sq<false, false>(mi, XtX, R, T1, stream);                       // XtX R
  vu_form_G_kernel<<<eb, 256, 0, stream>>>(G3, T1, W, alpha, mi, G);    // G
  dgemm_sq<true, false>(mi, G, G, M, stream);                           // G^T G
  dgemm_sq<false, false>(mi, M, basis, T2, stream);                     // (G^T G) basis
  dgemm_sq<true, false>(mi, basis, T2, Mp, stream);                     // basis^T (G^T G) basis
  XB_LAUNCH_CHECK();
  // eig_tol: relative off-diagonal norm at which the Jacobi sweeps may stop (0: the solver's own 3e-15); the
  // tensor-core iterations carry fp32-level noise in G anyway and pass 1e-9
  int rc = sym_eig_launch(Mp, m, ev, Vp, ework, info, eig_tol > 0.0 ? eig_tol * eig_tol : 1e-29, stream);
  if (rc) return rc;
  dgemm_sq<false, false>(mi, basis, Vp, V, stream);                     // eigenvectors of G^T G
  vu_scale_kernel<<<1, 256, 0, stream>>>(V, ev, mi, P, dsum);           // V diag(1/svals), delta
  dgemm_sq<false, false>(mi, G, P, T3, stream);                         // G V diag(1/svals)  (= U)
  dgemm_sq<false, true>(mi, T3, V, R, stream);                          // R = U V^T
  XB_CUDA(cudaMemcpyAsync(basis, V, (size_t)mm * sizeof(double), cudaMemcpyDeviceToDevice, stream));
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}
