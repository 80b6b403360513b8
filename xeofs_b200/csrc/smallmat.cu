// The k-column linear algebra around the streaming passes: Gram (fp64), Cholesky + triangular inverse,
// right-multiplication by a small matrix, Jacobi eigen-solver, sign-rule reductions.
// Together they stand in for sklearn's LU/QR normalizers and scipy.linalg.svd(B) inside
// sklearn.utils.extmath.randomized_svd (reference call site linalg/decomposer.py:141-146) and for
// get_deterministic_sign_multiplier (utils/xarray_utils.py:273-301).
#include <math.h>

#include "common.cuh"

namespace xb {

int env_int(const char* name, int dflt);  // project_tc.cu

// ------------------------------------------------------------------------------------------------
// Gram matrix in fp64 (CholeskyQR needs the Gram to cond^2 accuracy).  Persistent blocks sweep chunks of n; the
// chunk is converted to fp64 once on its way into shared memory.  The l x l result is cut into TI x TI register
// tiles (NT per side) and only the tiles on or above the diagonal are computed; with few tiles the rows of the
// chunk are dealt to KG thread groups.  8 x 8 tiles read 16 doubles for 64 fma, which keeps the kernel on the fp64
// pipe rather than on shared-memory bandwidth.  Global atomics: one per entry per group per block.
constexpr int GR_CHUNK = 32;

template <int TI, int NT, int KG, int SIDE>
__global__ void __launch_bounds__(256)
gram_kernel(const float* __restrict__ M, int64_t n, int l, int64_t ld, double* __restrict__ G) {
  constexpr int LP = TI * NT;
  constexpr int NTILES = NT * (NT + 1) / 2;
  static_assert(NTILES * KG <= 256, "tiles x groups must fit the block");
  __shared__ __align__(16) double sm[GR_CHUNK][LP + 2];  // [n_local][column]
  const int tid = threadIdx.x;
  // thread -> (upper-triangular tile (ti <= tj), row group kg)
  int ti = -1, tj = -1;
  const int kg = tid / NTILES;
  if (tid < NTILES * KG) {
    int t = tid - kg * NTILES, row = 0;
    while (t >= NT - row) { t -= NT - row; ++row; }
    ti = row;
    tj = row + t;
  }
  double acc[TI][TI];
#pragma unroll
  for (int a = 0; a < TI; ++a)
#pragma unroll
    for (int b = 0; b < TI; ++b) acc[a][b] = 0.0;

  const int64_t n_chunks = (n + GR_CHUNK - 1) / GR_CHUNK;
  // the next chunk travels from global memory into registers while the current one is multiplied
  constexpr int PER = GR_CHUNK * LP / 256;
  float nxt[PER];
  auto fetch = [&](int64_t ch) {
    const int64_t n0 = ch * GR_CHUNK;
    if (SIDE == 1) {
      // space-side: element (n, j) at M[j*ld + n]; 32 lanes read 32 consecutive n of one row j
      const int lane = tid & 31, w = tid >> 5;
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int j = w + 8 * i;
        nxt[i] = (ch < n_chunks && j < l && n0 + lane < n) ? M[(int64_t)j * ld + n0 + lane] : 0.f;
      }
    } else {
      // time-side: element (n, j) at M[n*ld + j]
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int idx = tid + 256 * i, r = idx / LP, j = idx % LP;
        nxt[i] = (ch < n_chunks && j < l && n0 + r < n) ? M[(n0 + r) * ld + j] : 0.f;
      }
    }
  };
  fetch(blockIdx.x);
  for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    __syncthreads();  // everyone is done with the previous chunk
    if (SIDE == 1) {
      const int lane = tid & 31, w = tid >> 5;
#pragma unroll
      for (int i = 0; i < PER; ++i) sm[lane][w + 8 * i] = (double)nxt[i];
    } else {
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int idx = tid + 256 * i;
        sm[idx / LP][idx % LP] = (double)nxt[i];
      }
    }
    __syncthreads();
    fetch(ch + gridDim.x);
    if (ti >= 0) {
#pragma unroll 2
      for (int r = kg; r < GR_CHUNK; r += KG) {
        double a[TI], b[TI];
#pragma unroll
        for (int x = 0; x < TI; ++x) {
          a[x] = sm[r][ti * TI + x];
          b[x] = sm[r][tj * TI + x];
        }
#pragma unroll
        for (int x = 0; x < TI; ++x)
#pragma unroll
          for (int y = 0; y < TI; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
      }
    }
  }
  if (ti >= 0) {
#pragma unroll
    for (int x = 0; x < TI; ++x)
#pragma unroll
      for (int y = 0; y < TI; ++y) {
        const int i = ti * TI + x, j = tj * TI + y;
        if (i < l && j < l) {
          if (ti != tj || j >= i) atomicAdd(&G[(int64_t)i * l + j], acc[x][y]);
          if (j > i) atomicAdd(&G[(int64_t)j * l + i], acc[x][y]);
        }
      }
  }
}

// ------------------------------------------------------------------------------------------------
// Cholesky G = R^T R and Rinv = R^-1, one block, fp64 (full l x l in dynamic smem for R, Rinv written to
// global).  l <= 128.  A column whose pivot has fallen below the fp32 noise floor of the matrix it was
// accumulated from (4 * FLT_EPSILON^2 * G[k][k]) is linearly dependent on the earlier ones to working
// precision: it is dropped from the basis (its column of Rinv is zero, so the orthonormalised matrix gets
// a zero column there) instead of being normalised into noise.  info[0] = number of dropped columns,
// info[1] = 1 if a NaN/Inf pivot was met (numpy.linalg.LinAlgError at the boundary).
__global__ void __launch_bounds__(256)
chol_inv_kernel(const double* __restrict__ G, int l, double* __restrict__ Rinv, int32_t* __restrict__ info) {
  extern __shared__ double sh[];  // A: l x (l+1); diag0[l]; dead[l]; dinv[l]
  // The upper triangle of A becomes R; the strict lower triangle then receives R^-1 transposed
  // (Rinv[i][j], i < j, lives at A[j][i]) and dinv its diagonal: the whole factorisation stays in shared memory.
  const int ldA = l + 1;
  double* A = sh;
  double* diag0 = A + (size_t)l * ldA;
  double* dead = diag0 + l;
  double* dinv = dead + l;
  __shared__ int n_dead, bad;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  if (tid == 0) { n_dead = 0; bad = 0; }
  for (int idx = tid; idx < l * l; idx += blockDim.x) {
    const int i = idx / l, j = idx - i * l;
    A[i * ldA + j] = G[idx];
  }
  __syncthreads();
  for (int i = tid; i < l; i += blockDim.x) { diag0[i] = A[i * ldA + i]; dead[i] = 0.0; }
  __syncthreads();
  // right-looking upper Cholesky: row k of R, then trailing update A[i][j] -= R[k][i] R[k][j]  (j >= i > k)
  for (int k = 0; k < l; ++k) {
    const double d = A[k * ldA + k];
    const bool finite = (d == d) && fabs(d) < 1e300;
    const bool drop = !finite || !(d > 5.7e-14 * diag0[k]) || !(d > 0.0);
    __syncthreads();
    if (drop) {
      if (tid == 0) { dead[k] = 1.0; n_dead += 1; if (!finite) bad = 1; }
      for (int j = k + tid; j < l; j += blockDim.x) A[k * ldA + j] = (j == k) ? 1.0 : 0.0;
      __syncthreads();
      continue;
    }
    const double rkk = sqrt(d), rinv = 1.0 / rkk;
    for (int j = k + tid; j < l; j += blockDim.x) A[k * ldA + j] = (j == k) ? rkk : A[k * ldA + j] * rinv;
    __syncthreads();
    for (int i = k + 1 + ty; i < l; i += 16) {
      const double rki = A[k * ldA + i];
      for (int j = i + tx; j < l; j += 16) A[i * ldA + j] = fma(-rki, A[k * ldA + j], A[i * ldA + j]);
    }
    __syncthreads();
  }
  if (tid == 0) { info[0] = n_dead; info[1] = bad; }
  // inverse of upper-triangular R, row by row from the bottom:
  //   Rinv[i][i] = 1/R[i][i];  Rinv[i][j] = -(sum_{k=i+1..j} R[i][k] Rinv[k][j]) / R[i][i]   (j > i)
  for (int i = tid; i < l; i += blockDim.x) dinv[i] = dead[i] != 0.0 ? 0.0 : 1.0 / A[i * ldA + i];
  __syncthreads();
  for (int i = l - 2; i >= 0; --i) {
    const double ri = 1.0 / A[i * ldA + i];
    for (int j = i + 1 + tid; j < l; j += blockDim.x) {
      double a0 = 0.0, a1 = 0.0;
      int k = i + 1;
      for (; k + 1 < j; k += 2) {  // Rinv[k][j] (k < j) is stored at A[j][k]
        a0 = fma(A[i * ldA + k], A[j * ldA + k], a0);
        a1 = fma(A[i * ldA + k + 1], A[j * ldA + k + 1], a1);
      }
      if (k < j) a0 = fma(A[i * ldA + k], A[j * ldA + k], a0);
      a0 = fma(A[i * ldA + j], dinv[j], a0);
      A[j * ldA + i] = -(a0 + a1) * ri;
    }
    __syncthreads();
  }
  for (int idx = tid; idx < l * l; idx += blockDim.x) {
    const int i = idx / l, j = idx - i * l;
    Rinv[idx] = (j > i) ? A[j * ldA + i] : (j == i ? dinv[i] : 0.0);
  }
}

// ------------------------------------------------------------------------------------------------
// Out(n, j') = sum_j In(n, j) Mat[j, j'] colscale[j'].  Block: 64 n x all j' (<= 128), 256 threads, thread
// tile 4 n x TJ j'.  fp32 inputs, fp64 matrix rounded to fp32 for the space-side (S x l x k flops), fp64
// accumulation on the (tiny) time side.
constexpr int AP_BN = 64;

template <int SIDE_IN, int SIDE_OUT, typename AccT>
__global__ void __launch_bounds__(256)
apply_kernel(const float* __restrict__ In, int64_t n, int l, int64_t ld_in, const double* __restrict__ Mat,
             int64_t ldm, int k, const double* __restrict__ colscale, float* __restrict__ Out, int64_t ld_out,
             int kp_out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  // layout: Ms[l][KP] (AccT), Is[l][AP_BN + 4] (float)
  const int KP = (k + 15) / 16 * 16;
  AccT* Ms = reinterpret_cast<AccT*>(smraw);
  float* Is = reinterpret_cast<float*>(smraw + (size_t)l * KP * sizeof(AccT));
  constexpr int ISL = AP_BN + 4;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < l * KP; idx += 256) {
    const int j = idx / KP, jp = idx % KP;
    double v = 0.0;
    if (jp < k) v = Mat[(int64_t)j * ldm + jp] * (colscale ? colscale[jp] : 1.0);
    Ms[idx] = (AccT)v;
  }
  const int tn = tid & 15;   // 4 n each -> 64 n
  const int tjg = tid >> 4;  // 16 groups over j'
  const int n_jt = (KP + 15) / 16;  // j' per thread (KP/16), <= 8

  for (int64_t blk = blockIdx.x; blk * AP_BN < n; blk += gridDim.x) {
    const int64_t n0 = blk * AP_BN;
    __syncthreads();
    if (SIDE_IN == 1) {
      const int lane = tid & 31, w = tid >> 5;
      for (int j = w; j < l; j += 8) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int64_t nn = n0 + lane + 32 * h;
          Is[j * ISL + lane + 32 * h] = (nn < n) ? In[(int64_t)j * ld_in + nn] : 0.f;
        }
      }
    } else {
      for (int idx = tid; idx < AP_BN * l; idx += 256) {
        const int r = idx / l, j = idx % l;
        Is[j * ISL + r] = (n0 + r < n) ? In[(n0 + r) * ld_in + j] : 0.f;
      }
    }
    __syncthreads();
    AccT acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = (AccT)0;
    for (int j = 0; j < l; ++j) {
      const float4 iv = *reinterpret_cast<const float4*>(&Is[j * ISL + 4 * tn]);
      const AccT ia[4] = {(AccT)iv.x, (AccT)iv.y, (AccT)iv.z, (AccT)iv.w};
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        if (b < n_jt) {
          const AccT m = Ms[j * KP + tjg + 16 * b];
#pragma unroll
          for (int a = 0; a < 4; ++a) acc[a][b] = fma(ia[a], m, acc[a][b]);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      if (b >= n_jt) continue;
      const int jp = tjg + 16 * b;
      if (jp >= kp_out) continue;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int64_t nn = n0 + 4 * tn + a;
        if (nn >= n) continue;
        const float v = (jp < k) ? (float)acc[a][b] : 0.f;
        if (SIDE_OUT == 1) Out[(int64_t)jp * ld_out + nn] = v;
        else Out[nn * ld_out + jp] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The space-side Gram matrix (M = l rows of n contiguous values) on the fp64 tensor-core path: mma.sync m8n8k4,
// IEEE fp64 FMAs as in gram_kernel.  A chunk of 64 values per turn, converted to fp64 on its way into shared memory
// ([n][8 MT + 4] doubles: the fragment loads of a half-warp fall into 16 different bank pairs); the l x l result is
// cut into 8 x 8 tiles, only those on or above the diagonal are computed, and warp w owns the tile rows w and
// MT - 1 - w (MT + 1 tiles each), accumulators in registers over all the turns of the CTA.
constexpr int GM_CH = 64;

__device__ __forceinline__ void dmma884_g(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int MT>
__global__ void __launch_bounds__(256)
gram_mma_kernel(const float* __restrict__ M, int64_t n, int l, int64_t ld, double* __restrict__ G) {
  constexpr int MP = 8 * MT, LDP = MP + 4;
  extern __shared__ __align__(16) unsigned char gm_raw[];
  double* Ms = reinterpret_cast<double*>(gm_raw);  // [GM_CH][LDP]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lc = lane & 3;
  // MT <= 8: four warps hold all the tile rows, the other four take every second K step of the same rows
  constexpr int KSPLIT = MT <= 8 ? 2 : 1;
  const int wq = KSPLIT == 2 ? (warp & 3) : warp, kpar = KSPLIT == 2 ? (warp >> 2) : 0;
  const int mt0 = wq, mt1 = MT - 1 - wq;  // the two tile rows of this warp (the same one in the middle of an odd MT)
  const bool has0 = mt0 <= mt1, has1 = mt0 < mt1;
  double acc[2][MT][2];
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int b = 0; b < MT; ++b) acc[x][b][0] = acc[x][b][1] = 0.0;

  const int64_t n_chunks = (n + GM_CH - 1) / GM_CH;
  for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const int64_t s0 = ch * GM_CH;
    __syncthreads();
    {
      const int sl = tid & 63, w4 = tid >> 6;
      const int64_t s = s0 + sl;
      for (int j = w4; j < MP; j += 4) Ms[sl * LDP + j] = (j < l && s < n) ? (double)M[(int64_t)j * ld + s] : 0.0;
    }
    __syncthreads();
    if (has0) {
#pragma unroll 2
      for (int k = kpar; k < GM_CH / 4; k += KSPLIT) {
        const double* row = Ms + (4 * k + lc) * LDP + lr;
        const double a0 = row[8 * mt0], a1 = row[8 * mt1];
#pragma unroll
        for (int b = 0; b < MT; ++b) {
          if (b < mt0) continue;
          const double fb = row[8 * b];
          dmma884_g(acc[0][b][0], acc[0][b][1], a0, fb);
          if (has1 && b >= mt1) dmma884_g(acc[1][b][0], acc[1][b][1], a1, fb);
        }
      }
    }
  }
  // element e of tile (mt, b): G[8 mt + lr][8 b + 2 lc + e], mirrored below the diagonal
#pragma unroll
  for (int x = 0; x < 2; ++x) {
    if (x == 0 ? !has0 : !has1) continue;
    const int mt = x == 0 ? mt0 : mt1;
    const int i = 8 * mt + lr;
#pragma unroll
    for (int b = 0; b < MT; ++b) {
      if (b < mt) continue;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 8 * b + 2 * lc + e;
        if (i < l && j < l) {
          if (b > mt || j >= i) atomicAdd(&G[(int64_t)i * l + j], acc[x][b][e]);
          if (b > mt || j > i) atomicAdd(&G[(int64_t)j * l + i], acc[x][b][e]);
        }
      }
    }
  }
}

template <int MT>
static int launch_gram_mma(const float* M, int64_t n, int l, int64_t ld, double* G, cudaStream_t stream) {
  constexpr int LDP = 8 * MT + 4;
  const size_t smem = (size_t)GM_CH * LDP * sizeof(double);
  XB_CUDA(cudaFuncSetAttribute(gram_mma_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks = (int)imin(ceil_div(n, GM_CH), 2 * (int64_t)num_sms());
  gram_mma_kernel<MT><<<blocks, 256, smem, stream>>>(M, n, l, ld, G);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

// ------------------------------------------------------------------------------------------------
// Symmetric eigen-decomposition by parallel cyclic Jacobi (round-robin pairing), one block, fp64.
// A (l x l) lives in dynamic smem; the eigenvector matrix is accumulated TRANSPOSED (row p = eigenvector p) in
// shared memory when both fit (VS: l <= 118), else in `work` (global), and written out as columns at the end, sorted
// by descending eigenvalue.  Every round is two block-wide phases (rotations of the rows, then of the columns); the
// seating of the round-robin tournament is computed, not stored.
constexpr int EIG_THREADS = 1024;

// player at seat q (0 .. le-1) in round r (< le-1) of the circle method: seat 0 keeps player 0, the others rotate
__device__ __forceinline__ int eig_seat(int q, int r, int le) {
  if (q == 0) return 0;
  int x = q - 1 + r;
  if (x >= le - 1) x -= le - 1;
  return 1 + x;
}
// the two players of pair q in round r, smaller index first
__device__ __forceinline__ void eig_pair(int q, int r, int le, int& p, int& s) {
  p = eig_seat(q, r, le);
  s = eig_seat(le - 1 - q, r, le);
  if (p > s) { const int t = p; p = s; s = t; }
}

template <bool VS>
__global__ void __launch_bounds__(EIG_THREADS)
sym_eig_kernel(const double* __restrict__ G, int l, double* __restrict__ evals, double* __restrict__ evecs,
               double* __restrict__ work, int32_t* __restrict__ info, double tol2) {
  extern __shared__ double sh[];
  const int le = (l + 1) & ~1;  // even size for the tournament
  const int ldA = le + 1;
  double* A = sh;                      // le x ldA
  double* cs = A + (size_t)le * ldA;   // le/2 cosines, le/2 sines
  int* dest = reinterpret_cast<int*>(cs + le);  // le ints: output column of every eigenvector
  int* pr = dest + le;                          // le ints: the two players of every pair in this round (p < r)
  double* Vt = VS ? reinterpret_cast<double*>(pr + le) : work;  // le x le
  __shared__ double offnorm;
  __shared__ double diagnorm;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int half = le / 2;
  // 2 x 2 blocks (row pair qa, column pair qb >= qa) of this thread: fixed for the whole kernel (EIG_THREADS threads,
  // half <= 64: at most 3), packed qa << 8 | qb.  Only the blocks on or above the diagonal are computed; their
  // transposes are stored along (A stays exactly symmetric, and half of the loads go away: the kernel is bound by
  // its shared-memory traffic)
  constexpr int NBLK = 3;
  int blk[NBLK];
#pragma unroll
  for (int i = 0; i < NBLK; ++i) {
    const int b = tid + i * nt;
    blk[i] = -1;
    if (b < half * (half + 1) / 2) {
      int qa = 0, rem = b;
      while (rem >= half - qa) { rem -= half - qa; ++qa; }
      blk[i] = qa << 8 | (qa + rem);
    }
  }
  // eigenvector rows: tpq threads share the two rows of a pair
  const int tpq = nt / half > 0 ? nt / half : 1;
  const int vq = tid / tpq, vsub = tid - vq * tpq;

  for (int i = tid >> 5; i < le; i += nt >> 5)
    for (int j = tid & 31; j < le; j += 32) {
      double v = 0.0;
      if (i < l && j < l) v = 0.5 * (G[(int64_t)i * l + j] + G[(int64_t)j * l + i]);
      A[i * ldA + j] = v;
      Vt[i * le + j] = (i == j) ? 1.0 : 0.0;
    }
  __syncthreads();

  int sweeps_done = 0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    // convergence test: off-diagonal Frobenius norm vs diagonal
    if (tid == 0) { offnorm = 0.0; diagnorm = 0.0; }
    __syncthreads();
    double off = 0.0, dg = 0.0;
    for (int i = tid >> 5; i < le; i += nt >> 5)
      for (int j = tid & 31; j < le; j += 32) {
        const double v = A[i * ldA + j];
        if (i == j) dg += v * v; else off += v * v;
      }
    off = warp_sum(off); dg = warp_sum(dg);
    if ((tid & 31) == 0) { atomicAdd(&offnorm, off); atomicAdd(&diagnorm, dg); }
    __syncthreads();
    const bool done = offnorm <= tol2 * diagnorm || diagnorm == 0.0;
    __syncthreads();
    if (done) break;
    sweeps_done = sweep + 1;

    for (int round = 0; round < le - 1; ++round) {
      // rotation of pair q: zeroes A[p][r]
      if (tid < half) {
        int p, r;
        eig_pair(tid, round, le, p, r);
        const double apq = A[p * ldA + r];
        double c = 1.0, sn = 0.0;
        if (fabs(apq) > 1e-300) {
          // fp64 division and square root are long software sequences and this is the serial part of every round: the
          // tangent is taken in fp32 (the rotation then leaves ~1e-7 |apq| behind instead of 0, which the next sweep
          // removes), while c and s are made orthonormal to fp64 accuracy from it
          const double app = A[p * ldA + p], aqq = A[r * ldA + r];
          const double d = aqq - app;
          int e;  // a power-of-two scale keeps both inside the fp32 range
          (void)frexp(fmax(fabs(d), fabs(apq)), &e);
          const float tau = (float)ldexp(d, -e) / (2.f * (float)ldexp(apq, -e));
          const float t32 = (tau >= 0.f ? 1.f : -1.f) / (fabsf(tau) + sqrtf(1.f + tau * tau));
          const double t = isfinite(t32) ? (double)t32 : 0.0;
          c = rsqrt(1.0 + t * t);
          sn = t * c;
        }
        cs[tid] = c; cs[half + tid] = sn;
        pr[tid] = p; pr[half + tid] = r;
      }
      __syncthreads();
      // A <- J^T A J, one 2 x 2 block (row pair qa, column pair qb >= qa) at a time: every element is written once
#pragma unroll
      for (int i = 0; i < NBLK; ++i) {
        if (blk[i] < 0) continue;
        const int qa = blk[i] >> 8, qb = blk[i] & 255;
        int pa, ra, pb, rb;  // (recomputed, not read back from pr: the kernel is bound by shared-memory traffic)
        eig_pair(qa, round, le, pa, ra);
        eig_pair(qb, round, le, pb, rb);
        const double ca = cs[qa], sa = cs[half + qa], cb = cs[qb], sb = cs[half + qb];
        const double a00 = A[pa * ldA + pb], a01 = A[pa * ldA + rb], a10 = A[ra * ldA + pb], a11 = A[ra * ldA + rb];
        const double b00 = ca * a00 - sa * a10, b01 = ca * a01 - sa * a11;
        const double b10 = sa * a00 + ca * a10, b11 = sa * a01 + ca * a11;
        const double n00 = cb * b00 - sb * b01, n01 = sb * b00 + cb * b01;
        const double n10 = cb * b10 - sb * b11, n11 = sb * b10 + cb * b11;
        A[pa * ldA + pb] = n00;
        A[pa * ldA + rb] = n01;
        A[ra * ldA + pb] = n10;
        A[ra * ldA + rb] = n11;
        if (qa != qb) {
          A[pb * ldA + pa] = n00;
          A[rb * ldA + pa] = n01;
          A[pb * ldA + ra] = n10;
          A[rb * ldA + ra] = n11;
        }
      }
      // eigenvector accumulator: rows p, r <- J^T rows
      for (int q = vq; q < half; q += (nt + tpq - 1) / tpq) {
        const int p = pr[q], r = pr[half + q];
        const double c = cs[q], sn = cs[half + q];
        for (int col = vsub; col < le; col += tpq) {
          if (VS) {
            const double vp = Vt[p * le + col], vr = Vt[r * le + col];
            Vt[p * le + col] = c * vp - sn * vr;
            Vt[r * le + col] = sn * vp + c * vr;
          } else {
            const double vp = __ldcg(&Vt[p * le + col]), vr = __ldcg(&Vt[r * le + col]);
            __stcg(&Vt[p * le + col], c * vp - sn * vr);
            __stcg(&Vt[r * le + col], sn * vp + c * vr);
          }
        }
      }
      __syncthreads();
    }
  }
  // sort eigenvalues descending (l <= 128: rank by counting), write outputs
  for (int i = tid; i < l; i += nt) {
    const double vi = A[i * ldA + i];
    int rank = 0;
    for (int j = 0; j < l; ++j) {
      const double vj = A[j * ldA + j];
      rank += (vj > vi) || (vj == vi && j < i);
    }
    evals[rank] = vi;
    dest[i] = rank;
  }
  __syncthreads();
  for (int idx = tid; idx < l * l; idx += nt) {
    const int i = idx / l, comp = idx % l;  // eigenvector i, component comp
    evecs[(int64_t)comp * l + dest[i]] = VS ? Vt[i * le + comp] : __ldcg(&Vt[i * le + comp]);
  }
  if (tid == 0) info[0] = sweeps_done;
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  // ordered-int trick, valid for any finite floats
  if (v >= 0) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  if (v >= 0) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void minmax_init_kernel(float* vmax, float* vmin, int k) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < k) { vmax[i] = -INFINITY; vmin[i] = INFINITY; }
}

__global__ void __launch_bounds__(256)
row_minmax_kernel(const float* __restrict__ Vt, int64_t n, int64_t ld, float* __restrict__ vmax, float* __restrict__ vmin) {
  const float* row = Vt + (int64_t)blockIdx.y * ld;
  float mx = -INFINITY, mn = INFINITY;
  auto take = [&](float v) {
    if (v == v) { mx = fmaxf(mx, v); mn = fminf(mn, v); }
  };
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((((uintptr_t)row) & 15) == 0) {
    // 16-byte loads, four in flight per thread
    const float4* row4 = reinterpret_cast<const float4*>(row);
    const int64_t n4 = n >> 2;
    for (; i + 3 * stride < n4; i += 4 * stride) {
      const float4 a = __ldg(row4 + i), b = __ldg(row4 + i + stride), c = __ldg(row4 + i + 2 * stride),
                   d = __ldg(row4 + i + 3 * stride);
      take(a.x); take(a.y); take(a.z); take(a.w); take(b.x); take(b.y); take(b.z); take(b.w);
      take(c.x); take(c.y); take(c.z); take(c.w); take(d.x); take(d.y); take(d.z); take(d.w);
    }
    for (; i < n4; i += stride) {
      const float4 a = __ldg(row4 + i);
      take(a.x); take(a.y); take(a.z); take(a.w);
    }
    for (int64_t j = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) take(row[j]);
  } else {
    for (; i < n; i += stride) take(row[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (mx > -INFINITY) atomic_max_float(&vmax[blockIdx.y], mx);
    if (mn < INFINITY) atomic_min_float(&vmin[blockIdx.y], mn);
  }
}

__global__ void finish_components_kernel(float* __restrict__ Vt, int64_t n, int64_t ld, const float* __restrict__ sign,
                                         const uint8_t* __restrict__ valid) {
  float* row = Vt + (int64_t)blockIdx.y * ld;
  const float sg = sign ? sign[blockIdx.y] : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const bool ok = valid ? (valid[i] != 0) : true;
    row[i] = ok ? sg * row[i] : nanf("");
  }
}

}  // namespace xb

using namespace xb;

extern "C" int xeofs_b200_gram(const float* M, int64_t n, int64_t l, int64_t ld, int side, double* G, int accumulate,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(M && G && n > 0 && l > 0 && l <= 128, "gram: bad arguments (l=%lld must be in 1..128)", (long long)l);
  XB_CHECK_ARG(side == 0 || side == 1, "gram: side must be 0 (time-side) or 1 (space-side)");
  if (!accumulate) XB_CUDA(cudaMemsetAsync(G, 0, (size_t)l * l * sizeof(double), stream));
  if (side == 1 && l > 32 && n >= 4096 && env_int("XEOFS_GRAM_MMA", 1))
    return l <= 64    ? launch_gram_mma<8>(M, n, (int)l, ld, G, stream)
           : l <= 104 ? launch_gram_mma<13>(M, n, (int)l, ld, G, stream)
                      : launch_gram_mma<16>(M, n, (int)l, ld, G, stream);
  const int64_t chunks = ceil_div(n, GR_CHUNK);
  const int blocks = (int)imin(chunks, (l <= 64 ? 6 : 3) * (int64_t)num_sms());
#define XB_GRAM(TI, NT, KG)                                                                      \
  if (side == 0) gram_kernel<TI, NT, KG, 0><<<blocks, 256, 0, stream>>>(M, n, (int)l, ld, G);           \
  else gram_kernel<TI, NT, KG, 1><<<blocks, 256, 0, stream>>>(M, n, (int)l, ld, G)
  if (l <= 16) { XB_GRAM(2, 8, 4); }
  else if (l <= 32) { XB_GRAM(4, 8, 4); }
  else if (l <= 64) { XB_GRAM(4, 16, 1); }
  else { XB_GRAM(8, 16, 1); }
#undef XB_GRAM
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

extern "C" int xeofs_b200_chol_inv(const double* G, int64_t l, double* Rinv, int32_t* info, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(G && Rinv && info && l > 0 && l <= 128, "chol_inv: bad arguments (l=%lld must be in 1..128)", (long long)l);
  const size_t smem = ((size_t)l * (l + 1) + 3 * (size_t)l) * sizeof(double);
  XB_CUDA(cudaFuncSetAttribute(chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  chol_inv_kernel<<<1, 256, smem, stream>>>(G, (int)l, Rinv, info);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

extern "C" int xeofs_b200_apply(const float* In, int64_t n, int64_t l, int64_t ld_in, int side, const double* Mat,
                                int64_t ldm, int64_t k, const double* colscale, float* Out, int64_t ld_out,
                                void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(In && Mat && Out && n > 0, "apply: null pointer");
  XB_CHECK_ARG(l > 0 && l <= 128 && k > 0 && k <= 128, "apply: l=%lld, k=%lld must be in 1..128", (long long)l, (long long)k);
  XB_CHECK_ARG(side == 0 || side == 1, "apply: side must be 0 or 1");
  const int KP = (int)lpad(k);
  const int blocks = (int)imin(ceil_div(n, AP_BN), 4 * (int64_t)num_sms());
  if (side == 1) {
    // space-side: fp32 accumulation (S*l*k flops), output keeps lp rows (pad rows written as zero)
    const size_t smem = (size_t)l * KP * sizeof(float) + (size_t)l * (AP_BN + 4) * sizeof(float);
    XB_CUDA(cudaFuncSetAttribute(apply_kernel<1, 1, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    apply_kernel<1, 1, float><<<blocks, 256, smem, stream>>>(In, n, (int)l, ld_in, Mat, ldm, (int)k, colscale, Out,
                                                            ld_out, KP);
  } else {
    const size_t smem = (size_t)l * KP * sizeof(double) + (size_t)l * (AP_BN + 4) * sizeof(float);
    XB_CUDA(cudaFuncSetAttribute(apply_kernel<0, 0, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int kp_out = (int)imin(KP, ld_out);
    apply_kernel<0, 0, double><<<blocks, 256, smem, stream>>>(In, n, (int)l, ld_in, Mat, ldm, (int)k, colscale, Out,
                                                             ld_out, kp_out);
  }
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

// tol2: the sweeps end when the squared off-diagonal Frobenius norm falls below tol2 x the squared diagonal norm
namespace xb {
int sym_eig_launch(const double* G, int64_t l, double* evals, double* evecs, double* work, int32_t* info, double tol2,
                   cudaStream_t stream);
}
extern "C" int xeofs_b200_sym_eig(const double* G, int64_t l, double* evals, double* evecs, double* work,
                                  int32_t* info, void* stream_) {
  return xb::sym_eig_launch(G, l, evals, evecs, work, info, 1e-29, (cudaStream_t)stream_);
}
int xb::sym_eig_launch(const double* G, int64_t l, double* evals, double* evecs, double* work, int32_t* info,
                       double tol2, cudaStream_t stream) {
  XB_CHECK_ARG(G && evals && evecs && work && info && l > 0 && l <= 128, "sym_eig: bad arguments (l=%lld must be in 1..128)", (long long)l);
  const int le = ((int)l + 1) & ~1;
  const size_t base = ((size_t)le * (le + 1) + le) * sizeof(double) + (size_t)2 * le * sizeof(int);
  const size_t with_v = base + (size_t)le * le * sizeof(double);
  const int threads = EIG_THREADS;  // the kernel's block table assumes half^2 <= 4 * threads
  if (with_v <= 227 * 1024) {
    XB_CUDA(cudaFuncSetAttribute(sym_eig_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)with_v));
    sym_eig_kernel<true><<<1, threads, with_v, stream>>>(G, (int)l, evals, evecs, work, info, tol2);
  } else {
    XB_CUDA(cudaFuncSetAttribute(sym_eig_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)base));
    sym_eig_kernel<false><<<1, threads, base, stream>>>(G, (int)l, evals, evecs, work, info, tol2);
  }
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

extern "C" int xeofs_b200_row_minmax(const float* Vt, int64_t k, int64_t n, int64_t ld, float* vmax, float* vmin,
                                     void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(Vt && vmax && vmin && k > 0 && k <= 65535 && n > 0, "row_minmax: bad arguments");
  minmax_init_kernel<<<(unsigned)ceil_div(k, 128), 128, 0, stream>>>(vmax, vmin, (int)k);
  XB_LAUNCH_CHECK();
  const int bx = (int)imin(ceil_div(n, 256 * 16), 2 * (int64_t)num_sms());
  row_minmax_kernel<<<dim3(bx > 0 ? bx : 1, (unsigned)k), 256, 0, stream>>>(Vt, n, ld, vmax, vmin);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

extern "C" int xeofs_b200_finish_components(float* Vt, int64_t k, int64_t n, int64_t ld, const float* sign,
                                            const uint8_t* valid, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(Vt && k > 0 && k <= 65535 && n > 0, "finish_components: bad arguments");
  const int bx = (int)imin(ceil_div(n, 256 * 4), 4 * (int64_t)num_sms());
  finish_components_kernel<<<dim3(bx > 0 ? bx : 1, (unsigned)k), 256, 0, stream>>>(Vt, n, ld, sign, valid);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}
