/*
 * xeofs_b200.h — C-ABI of libxeofs_b200.so: the B200 (sm_100a) kernels behind
 * xeofs.single.EOF.fit / xeofs.cross.MCA.fit / xeofs.single.EOFRotator.fit.
 *
 * The reference (xeofs v3.0.4) is pure Python and has no FFI; every entry point below therefore
 * replaces a span of Python/numpy/scikit-learn code, cited per function as file:line under
 * /root/reference/xeofs.  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer (cudaMalloc / torch.cuda tensor .data_ptr()) unless the
 *    parameter name ends in _host; the library never allocates persistent memory and never frees
 *    caller memory; scratch is the caller-provided workspace (query the size first);
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*), nothing synchronises;
 *  - return value: 0 = ok, <0 = error class (XEOFS_E_*); the message is thread-local, see
 *    xeofs_b200_last_error();
 *  - layouts (row-major, leading dimension in ELEMENTS):
 *      X      "field"           T x S   fp32, ldx >= S          (time x space, space contiguous)
 *      Yt     "space-side"      lp x S  fp32, ldy >= S          (mode-major: row j = column j of the S x l matrix)
 *      W / Z  "time-side"       T x lp  fp32, ldw >= lp
 *      lp = l rounded up to a multiple of 16; pad rows/columns are zero and stay zero;
 *  - per-feature vectors (length S): pivot (subtracted before tensor-core rounding), dscale
 *    (= valid * coslat * weight / std), ccorr (rank-1 correction (pivot - mean_eff) * dscale, may be NULL = 0).
 *    Effective matrix:  A[t,s] = (X[t,s] - pivot[s]) * dscale[s] + ccorr[s];  a NaN entry counts as 0 in the
 *    first term.  Samples that are NaN throughout (dropped by the Sanitizer, sanitizer.py:49-50,124) are named by
 *    row_valid[T] (uint8, NULL = every sample valid): such a row of A is zero altogether.
 */
#ifndef XEOFS_B200_H
#define XEOFS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XEOFS_OK 0
#define XEOFS_E_INVALID (-1)   /* bad argument (maps to ValueError)            */
#define XEOFS_E_CUDA (-2)      /* CUDA runtime / launch failure (RuntimeError) */
#define XEOFS_E_WORKSPACE (-3) /* workspace too small (ValueError)             */
#define XEOFS_E_UNSUPPORTED (-4)

/* algo selector for the two streaming products */
#define XEOFS_ALGO_AUTO 0
#define XEOFS_ALGO_SIMT 1   /* fp32 CUDA-core kernels (validation path, any alignment)          */
#define XEOFS_ALGO_TF32X1 2 /* tcgen05 kind::tf32, one product (power iterations)               */
#define XEOFS_ALGO_TF32X3 3 /* tcgen05 kind::tf32, hi/lo split, 3 products (~fp32 accuracy)     */
#define XEOFS_ALGO_AUTO_FAST 4 /* TF32X1 where the tcgen05 path applies, else SIMT (power iterations) */
#define XEOFS_ALGO_TF32X2 5 /* tcgen05 kind::tf32, field split hi/lo, the small operand (W / Yt) must hold TF32-exact
                               values (xeofs_b200_round_tf32): 2 products, ~fp32 accuracy; SIMT where tcgen05 does
                               not apply                                                                        */
/* OR-ed into `algo` of project_T: the caller vouches that the field holds no NaN / Inf at all (every feature and every
 * sample valid, row_valid NULL), so the tensor-core operand stage skips its per-value test.                          */
#define XEOFS_ALGO_FLAG_NO_NAN 0x100
#define XEOFS_ALGO_TF32X1F 7 /* as TF32X1R for a field that is ALREADY a TF32-rounded copy of the preprocessed matrix
                               (xeofs_b200_materialize with round_tf32, pivot 0, dscale 1): no rounding, no test  */
#define XEOFS_ALGO_TF32X1R 6 /* TF32X1 with both operands rounded to TF32 to nearest (unbiased sums; else as TF32X1) */

/* flags for xeofs_b200_scaling_finalize */
#define XEOFS_F_CENTER 1
#define XEOFS_F_STANDARDIZE 2

int xeofs_b200_version(void);
const char* xeofs_b200_last_error(void);
/* 1 if the running device is sm_100 and the tcgen05 kernels can be used */
int xeofs_b200_has_tcgen05(void);

/* ---- P1/P4/P5: one streaming pass of column statistics -------------------------------------------
 * Replaces Scaler.fit's X.mean / X.std (preprocessing/scaler.py:100-108), Sanitizer's three notnull
 * reductions (preprocessing/sanitizer.py:46-56) and feeds total_variance (utils/xarray_utils.py:236-253).
 * Outputs (caller zero-initialises nothing; the call clears them):
 *   shift[S]   fp32  the per-column shift used for the sums (first row, 0 where that is NaN)
 *   sum[S]     fp64  sum over non-NaN t of (x - shift)
 *   sumsq[S]   fp64  sum over non-NaN t of (x - shift)^2
 *   cnt[S]     i32   number of non-NaN samples
 *   row_nan[T] i32   number of NaN features in each sample                                            */
int xeofs_b200_col_stats(const float* X, int64_t T, int64_t S, int64_t ldx, float* shift, double* sum,
                         double* sumsq, int32_t* cnt, int32_t* row_nan, void* stream);

/* ---- Scaler.fit tail + what Scaler.transform needs (preprocessing/scaler.py:100-153) -------------
 * From the raw statistics: mean (fp32), std (ddof 0, clipped at FLT_EPSILON, fp32), valid mask, and the
 * three vectors the streaming kernels consume.  featw[S] = coslat*weights per feature (NULL = 1).
 * scalars_out[0] = total variance  sum_s dscale^2 * M2_s / (cnt_s - 1)   (utils/xarray_utils.py:236-253)
 * scalars_out[1] = number of valid features, scalars_out[2] = max cnt, scalars_out[3] = min cnt over valid.  */
int xeofs_b200_scaling_finalize(int64_t S, const float* shift, const double* sum, const double* sumsq,
                                const int32_t* cnt, const double* featw, int flags, float* mean, float* std,
                                uint8_t* valid, float* pivot, float* dscale, float* ccorr,
                                double* scalars_out, void* stream);

/* ---- D2: the tall-skinny products of the randomized range finder ---------------------------------
 * sklearn.utils.extmath.randomized_range_finder's  M @ Q  and  M.T @ Q  (called from
 * linalg/decomposer.py:141-146), and cpcca.py:204-205's X.Q, with the Scaler arithmetic
 * (scaler.py:146-153) folded into the operand load so X is read once per pass.
 *
 * project_S:  Yt[j,s] = sum_t A[t,s] * W[t,j]          (A^T W, output space-side,  lp x S)
 * project_T:  Z[t,j]  = sum_s A[t,s] * Yt[j,s]         (A Y,   output time-side,   T x lp)
 * l is the live column count, lp = round_up(l,16) the stored one.                                    */
int64_t xeofs_b200_project_workspace_bytes(int64_t T, int64_t S, int64_t l, int algo);
int xeofs_b200_project_S(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot,
                         const float* dscale, const float* ccorr, const uint8_t* row_valid, const float* W,
                         int64_t ldw, int64_t l,
                         float* Yt, int64_t ldy, void* workspace, int64_t workspace_bytes, int algo,
                         void* stream);
int xeofs_b200_project_T(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot,
                         const float* dscale, const float* ccorr, const uint8_t* row_valid, const float* Yt,
                         int64_t ldy, int64_t l,
                         float* Z, int64_t ldz, void* workspace, int64_t workspace_bytes, int algo,
                         void* stream);

/* ---- P1/P4/P5 + D2 fused: statistics and the first product of the range finder from ONE read of X --------------
 * What col_stats + scaling_finalize + project_S(TF32X1) give, for the case where M = A^T comes first (n_samples <
 * n_features): the first sample of each feature is the shift of the sums and the pivot of this pass, the Scaler
 * vectors are derived in the epilogue of the CTA that owns the feature and applied to its block of Yt.
 * Every sample is taken as present: if row_nan then names samples that are NaN throughout, Yt's rank-1 term is off
 * for un-centred data and the sketch rows do not line up with the reference's compacted matrix — redo project_S.
 * Returns XEOFS_E_UNSUPPORTED where the tcgen05 path does not apply (use the three separate calls).               */
int xeofs_b200_project_S_stats(const float* X, int64_t T, int64_t S, int64_t ldx, const double* featw, int flags,
                               const float* W, int64_t ldw, int64_t l, float* mean, float* std, uint8_t* valid,
                               float* pivot, float* dscale, float* ccorr, double* scalars_out, int32_t* row_nan,
                               float* Yt, int64_t ldy, void* workspace, int64_t workspace_bytes, void* stream);

/* In place: every value of the (rows x cols, leading dimension ld) matrix keeps only the bits a TF32 operand has
 * (the low 13 mantissa bits are cleared), so that TF32X2 products with it as the small operand are exact.        */
int xeofs_b200_round_tf32(float* M, int64_t rows, int64_t cols, int64_t ld, void* stream);

/* ---- D2: the k-column orthonormalisation (sklearn's LU / QR normalizers) as CholeskyQR -----------
 * gram:      G[i,j] (+)= sum_n M(n,i) M(n,j), fp64, l x l row-major.  side 0: M is time-side (n x ld),
 *            side 1: M is space-side (lp x ld, n along the contiguous axis).  accumulate=0 clears G first.
 * chol_inv:  G = R^T R (upper R);  Rinv = R^-1 (l x l fp64 row-major, upper).  A column whose pivot falls
 *            below the fp32 noise floor of G (4 eps32^2 G[k][k]) is linearly dependent to working precision and
 *            is dropped: its column of Rinv is zero.  info[0] = number of dropped columns, info[1] = 1 if a
 *            non-finite pivot was met (numpy.linalg.LinAlgError at the boundary, decomposer.py:265-270).
 * apply:     Out(n, j') = sum_j In(n, j) * Mat[j, j'] * colscale[j']   for j < l, j' < k; Mat is l x k fp64
 *            row-major (ldm), colscale NULL = 1.  side as in gram; In/Out may alias only if identical.    */
int xeofs_b200_gram(const float* M, int64_t n, int64_t l, int64_t ld, int side, double* G, int accumulate,
                    void* stream);
int xeofs_b200_chol_inv(const double* G, int64_t l, double* Rinv, int32_t* info, void* stream);
int xeofs_b200_apply(const float* In, int64_t n, int64_t l, int64_t ld_in, int side, const double* Mat,
                     int64_t ldm, int64_t k, const double* colscale, float* Out, int64_t ld_out, void* stream);

/* ---- D2/D3: the small SVD and the sign rule ------------------------------------------------------
 * sym_eig: cyclic Jacobi on the l x l Gram (replaces scipy.linalg.svd(B) inside randomized_svd via
 *          B B^T = U diag(s^2) U^T).  evals[l] descending, evecs l x l row-major with eigenvector i in
 *          COLUMN i.  l <= 128.  work: le*le doubles with le = l rounded up to even (the eigenvector accumulator when
 *          it does not fit shared memory next to the matrix, l > 118; (l+1)*(l+1) doubles always suffice).
 * row_minmax: per row of a (k x n) matrix the max and the min (utils/xarray_utils.py:273-301 needs both).
 * finish_components: Vt[m,s] = valid[s] ? sign[m] * Vt[m,s] : NaN   (decomposer.py:219-222 and
 *          sanitizer.py:128-153's reindex).                                                          */
int xeofs_b200_sym_eig(const double* G, int64_t l, double* evals, double* evecs, double* work, int32_t* info,
                       void* stream);
int xeofs_b200_row_minmax(const float* Vt, int64_t k, int64_t n, int64_t ld, float* vmax, float* vmin,
                          void* stream);
int xeofs_b200_finish_components(float* Vt, int64_t k, int64_t n, int64_t ld, const float* sign,
                                 const uint8_t* valid, void* stream);

/* ---- R1: one varimax/promax sweep over the loadings (linalg/_numpy/_rotation.py:57-62, 166-177) ----
 * L is space-side (mp x S, mode-major), rownorm[S] the Kaiser normaliser 1/(h+eps) (NULL = 1).
 * With B = diag(rownorm) L^T R (S x m) and Ln = diag(rownorm) L^T:
 *   Gout[i,j] (+)= sum_s Ln[s,i] * f(B[s,j]),   f(b) = (b c_j) |b c_j|^(power-1),  c = colscale (NULL = 1)
 *   Wout[j]   (+)= sum_s B[s,j]^2,              absmax[j] = max_s |B[s,j]|  (fp32, may be NULL)
 * varimax: power = 3, colscale NULL; the caller forms G = Gout - (gamma/S) (Ln^T Ln) R diag(W).
 * promax target regression: power = p, colscale = 1/absmax, caller forms R^T Gout = Z^T P.               */
int xeofs_b200_varimax_accumulate(const float* L, int64_t S, int64_t m, int64_t ld, const float* rownorm,
                                  const double* R, double power, const double* colscale, double* Gout,
                                  double* Wout, float* absmax, int accumulate, void* stream);
/* The varimax sweep (power = 3, colscale NULL, rownorm NULL: Ln already Kaiser-normalised) on the tensor cores:
 * both products of _rotation.py:166-170 (B = X R and X^H B^3) as tcgen05 kind::tf32 MMAs with hi/lo split operands
 * (~fp32 accuracy per product, fp64 accumulation across tiles), the tile of Ln read from HBM once.  Ln is space-side
 * with lpad(m) rows (pad rows zero).  products = 3: hi/lo split operands as described; products = 1: one product per
 * GEMM with operands rounded to TF32 (a third of the tensor work, ~1e-3 per term, averaging out over the features) for
 * the iterations in which the rotation is still far from converged.  Returns XEOFS_E_UNSUPPORTED where tcgen05 does
 * not apply (use xeofs_b200_varimax_accumulate).
 * `packed` (optional, NULL = none): the copy xeofs_b200_varimax_pack made of the same Ln — the loadings do not change
 * during a rotation, so they are rewritten once, tile by tile (64 features x lpad(m) modes), as the shared-memory
 * image the tensor core reads; the sweep then fetches every tile with one bulk copy and works on 64 features per
 * MMA instead of 32 (1.8 -> 1.1 ms per sweep at 4.1 M features x 100 modes).  xeofs_b200_varimax_pack_bytes returns
 * the size of that buffer (128-byte aligned), 0 where the packed kernel does not apply.                             */
int64_t xeofs_b200_varimax_workspace_bytes(int64_t S, int64_t m);
int64_t xeofs_b200_varimax_pack_bytes(int64_t S, int64_t m);
int xeofs_b200_varimax_pack(const float* Ln, int64_t S, int64_t m, int64_t ld, float* packed, int64_t packed_bytes,
                            void* stream);
int xeofs_b200_varimax_sweep(const float* Ln, const float* packed, int64_t S, int64_t m, int64_t ld, const double* R,
                             double* Gout, double* Wout, int accumulate, int products, void* workspace,
                             int64_t workspace_bytes, void* stream);
/* The m x m step of one varimax iteration (_rotation.py:170-175), on the device: with Gout / Wout of the sweep, XtX =
 * Ln^T Ln and alpha = gamma / n_rows,  G = Gout - alpha (XtX R) diag(Wout);  R <- U V^T of svd(G) (in place);
 * *dsum = sum(svals).  `basis` (m x m, in/out; identity before the first iteration) carries the eigenvectors of
 * G^T G from one iteration to the next — in that basis the matrix is nearly diagonal and the Jacobi solver ends early.
 * eig_tol: relative off-diagonal norm at which the Jacobi sweeps may stop (0 = the solver's own 3e-15).
 * m <= 128; all matrices fp64 row-major.                                                                            */
int64_t xeofs_b200_varimax_update_workspace_bytes(int64_t m);
int xeofs_b200_varimax_update(const double* Gout, const double* Wout, const double* XtX, double alpha, int64_t m,
                              double* R, double* basis, double* dsum, double eig_tol, void* workspace,
                              int64_t workspace_bytes, void* stream);
/* Kaiser norms (_rotation.py:155-160): h[s] = sqrt(sum_j L[j,s]^2); rownorm[s] = 1/(h+eps);
 * Ln[j,s] = L[j,s] * rownorm[s] (space-side, ldn >= S).  Any of h / rownorm / Ln may be NULL.             */
int xeofs_b200_col_norms(const float* L, int64_t S, int64_t m, int64_t ld, float* h, float* rownorm, float* Ln,
                         int64_t ldn, void* stream);

/* ---- (f1) EOF.inverse_transform (single/eof.py:134-156 + scaler.py:165-190 + sanitizer.py:128-153) ----------
 * out[t,s] = (sum_i scores[t,i] * Vt[modes[i], s] - ccorr[s]) / dscale[s] + pivot[s], NaN where valid[s] == 0.
 * scores T x lds (time-side), Vt space-side (rows = modes), modes[m] int32 row indices into Vt.              */
int xeofs_b200_reconstruct(const float* scores, int64_t T, int64_t lds, const float* Vt, int64_t S, int64_t ldv,
                           const int32_t* modes, int64_t m, const float* pivot, const float* dscale,
                           const float* ccorr, const uint8_t* valid, float* out, int64_t ldo, void* stream);

/* ---- M3: a block of preprocessed samples as a space-side matrix (cross/cpcca.py:991-1000) ------------------------
 * out[j, s] = A[t0 + j, s] for j < nrows, zero for nrows <= j < rows_out.  With project_T this yields 128 columns of
 * the T x T Gram matrices X X^T and Y Y^T at a time, from which sum |C|^2 = <X X^T, Y Y^T>_F / (n-1)^2.            */
int xeofs_b200_scaled_rows(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot,
                           const float* dscale, const float* ccorr, const uint8_t* row_valid, int64_t t0,
                           int64_t nrows, int64_t rows_out, float* out, int64_t ldo, void* stream);

/* ---- D2 beyond one 128-column block, and the small dense matrices of the cross models' PCA stage ----------------
 * (what the reference gets from LAPACK / BLAS through scipy.linalg.lu / qr / svd inside randomized_svd, call site
 * linalg/decomposer.py:141-146, and numpy in preprocessing/pca.py:94-131, whitener.py:111-133.)
 * dgemm:         C (m x n) = alpha op(A) op(B) + beta C, all fp64 row-major, op = transpose when trans_* != 0.
 * gram_wide:     G (l x l fp64 row-major) = M^T M for a k-column matrix with l up to 4096 columns (side as in gram),
 *                fp64 accumulation, one pair of 128-column blocks per launch.
 * sym_eig_wide:  eigen-decomposition of a symmetric positive semi-definite n x n matrix (n <= 4096) by multi-CTA
 *                one-sided Jacobi; evals descending, eigenvector i in COLUMN i of evecs (n x n row-major);
 *                info[0] = sweeps run (at most max_sweeps; <= 0: 16).  Workspace size from the query function.     */
int xeofs_b200_dgemm(int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                     int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, void* stream);
int xeofs_b200_gram_wide(const float* M, int64_t n, int64_t l, int64_t ld, int side, double* G, void* stream);
int64_t xeofs_b200_sym_eig_wide_workspace_bytes(int64_t n);
int xeofs_b200_sym_eig_wide(const double* G, int64_t n, double* evals, double* evecs, void* workspace,
                            int64_t workspace_bytes, int32_t* info, int max_sweeps, void* stream);

/* The preprocessed matrix itself, A[t,s] = (X[t,s] - pivot[s]) dscale[s] + ccorr[s] (NaN -> 0, all-NaN samples zero),
 * written out as a T x S fp32 matrix with `rows_out` >= T rows (the extra rows zero) — for the sample Gram matrices
 * behind the total squared covariance (cross/cpcca.py:991-1000), which stream the field ~T/256 times: one rounded copy
 * lets those passes run as plain moves (XEOFS_ALGO_TF32X1F).  round_tf32 != 0: values rounded to TF32, ties to even. */
int xeofs_b200_materialize(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot, const float* dscale,
                           const float* ccorr, const uint8_t* row_valid, int64_t rows_out, int round_tf32, float* out,
                           int64_t ldo, void* stream);

/* ---- D2: a half-precision copy of the preprocessed matrix for the power iterations ------------------------------
 * The q power iterations of the range finder (sklearn _randomized_range_finder; call site linalg/decomposer.py:141-146)
 * only have to keep the right subspace: this build runs them as single TF32 products, which read 11 significant bits
 * of every operand.  fp16 holds those 11 bits at half the bytes: the first project_T pass of a fit also writes
 *     A16[t,s] = fp16( (X[t,s] - pivot[s]) e16[s] ),   e16 = dscale c,  c = 2^round(log2(512 / (|dscale| std)))
 * (round to nearest, NaN -> 0, |value| > 65504 = 127 sigma saturated) and the remaining power-iteration passes stream
 * that copy through kind::f16 products — half the HBM bytes per pass.  The two passes that decide the singular values
 * (XEOFS_ALGO_TF32X2 / TF32X3) always read the fp32 field.  Only for centred fields with every sample present.
 * h16_scales:        e16[s], ic16[s] = 1/c[s] from the Scaler vectors (0 for dropped features).
 * project_T_h16copy: xeofs_b200_project_T(XEOFS_ALGO_TF32X1) that also writes A16 (T x ldc halves, ldc >= S rounded
 *                    up to 128, ldc % 8 == 0); no_nan as XEOFS_ALGO_FLAG_NO_NAN.
 * project_S16 / project_T16: Yt = A^T W / Z = A Y with A[t,s] = A16[t,s] ic16[s] (+ cc16[s], see below) (same layouts
 *                    and workspace size as project_S / project_T with XEOFS_ALGO_TF32X1).                            */
int xeofs_b200_h16_scales(const float* dscale, const float* std, int64_t S, float* e16, float* ic16, void* stream);
int xeofs_b200_project_T_h16copy(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot,
                                 const float* dscale, const float* Yt, int64_t ldy, int64_t l, float* Z, int64_t ldz,
                                 void* workspace, int64_t workspace_bytes, int no_nan, const float* e16, void* copy16,
                                 int64_t ldc, void* stream);
int xeofs_b200_project_S16(const void* A16, int64_t T, int64_t S, int64_t ldc, const float* ic16, const float* cc16,
                           const float* W, int64_t ldw, int64_t l, float* Yt, int64_t ldy, void* workspace,
                           int64_t workspace_bytes, void* stream);
int xeofs_b200_project_T16(const void* A16, int64_t T, int64_t S, int64_t ldc, const float* ic16, const float* cc16,
                           const float* Yt, int64_t ldy, int64_t l, float* Z, int64_t ldz, void* workspace,
                           int64_t workspace_bytes, void* stream);
/* The statistics pass itself can write the copy (then the first project_T pass is already a half-precision one): the
 * mean is not known yet, so the copy holds the field shifted by its first sample,
 *     A16[t,s] = fp16( (X[t,s] - X[0,s]) c0[s] ),   c0 = the power of two that puts the largest deviation met in 64
 *     samples spread over the record at 512 (deviations 128 times larger still fit),
 * and the fitted matrix is A[t,s] = A16[t,s] ic16[s] + cc16[s] with ic16 = dscale / c0 and the rank-1 term
 * cc16 = (X[0,s] - mean[s]) dscale[s], both written by the pass.  project_S16 / project_T16 take cc16 (NULL for a
 * centred copy written by project_T_h16copy).  Same arguments as xeofs_b200_project_S_stats plus the copy.       */
int xeofs_b200_project_S_stats_h16copy(const float* X, int64_t T, int64_t S, int64_t ldx, const double* featw,
                                       int flags, const float* W, int64_t ldw, int64_t l, float* mean, float* std,
                                       uint8_t* valid, float* pivot, float* dscale, float* ccorr, double* scalars_out,
                                       int32_t* row_nan, float* Yt, int64_t ldy, void* workspace,
                                       int64_t workspace_bytes, void* copy16, int64_t ldc, float* c0, float* ic16,
                                       float* cc16, void* stream);

/* ---- M3: the sample Gram matrix A A^T of a preprocessed field as a plain tcgen05 GEMM (kind::f16 on a bf16 copy) ----
 * materialize_bf16: the preprocessed matrix (as xeofs_b200_materialize) rounded to bf16, rows_out x cols_out with zero
 *                   padding (rows_out a multiple of 256, cols_out a multiple of 64 for gram_rows_bf16).
 * gram_rows_bf16:   G (T_pad x ldg fp32) <- the block-lower triangle (t >= 256 floor(t'/256)) of A A^T, both operands
 *                   TMA-fed row tiles of the bf16 copy, fp32 accumulation in TMEM over at most 512 instructions per
 *                   accumulator, deterministic second-stage sum.  For sums over >= 65 536 features (the bf16 rounding
 *                   errors average out to ~5e-6 relative): the total squared covariance of cross/cpcca.py:991-1000 as
 *                   <X X^T, Y Y^T>_F / (n-1)^2.                                                                     */
int xeofs_b200_materialize_bf16(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot,
                                const float* dscale, const float* ccorr, const uint8_t* row_valid, int64_t rows_out,
                                int64_t cols_out, void* out, int64_t ldo, void* stream);
int64_t xeofs_b200_gram_rows_bf16_workspace_bytes(int64_t T_pad, int64_t S_pad);
int xeofs_b200_gram_rows_bf16(const void* A, int64_t T_pad, int64_t S_pad, int64_t ld, float* G, int64_t ldg,
                              void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XEOFS_B200_H */
