// fp32 CUDA-core versions of the two streaming products (XEOFS_ALGO_SIMT): the validation path for the
// tcgen05 kernels and the fallback for unaligned fields.  Same contract as project_tc.cu.
//   project_S:  Yt[j,s] = dscale[s] * sum_t (X[t,s]-pivot[s]) W[t,j]  +  ccorr[s] * sum_t W[t,j]
//   project_T:  Z[t,j]  = sum_s (X[t,s]-pivot[s]) dscale[s] Yt[j,s]   +  sum_s ccorr[s] Yt[j,s]
#include "common.cuh"

namespace xb {

// ------------------------------------------------------------------------------------------------
// column sums of W (T x lp) -> out[lp]   (needed for the rank-1 correction of project_S)
__global__ void colsum_kernel(const float* __restrict__ W, int64_t T, int64_t ldw, int lp,
                              const uint8_t* __restrict__ row_valid, float* __restrict__ out) {
  // one block per 32 columns, 8 warps stride over rows
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  const int w = threadIdx.x >> 5;
  double acc = 0;
  if (j < lp)
    for (int64_t t = w; t < T; t += 8)
      if (!row_valid || row_valid[t]) acc += (double)W[t * ldw + j];
  __shared__ double sh[8][32];
  sh[w][threadIdx.x & 31] = acc;
  __syncthreads();
  if (w == 0 && j < lp) {
    double a = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) a += sh[i][threadIdx.x];
    out[j] = (float)a;
  }
}

// r[j] = sum_s ccorr[s] * Yt[j,s]  (rank-1 correction of project_T), one block per row j
__global__ void ccorr_dot_kernel(const float* __restrict__ Yt, int64_t S, int64_t ldy, const float* __restrict__ ccorr,
                                 float* __restrict__ out) {
  const float* row = Yt + (int64_t)blockIdx.x * ldy;
  double acc = 0;
  for (int64_t s = threadIdx.x; s < S; s += blockDim.x) acc += (double)ccorr[s] * (double)row[s];
  acc = warp_sum(acc);
  __shared__ double sh[32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double a = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    a = warp_sum(a);
    if (threadIdx.x == 0) out[blockIdx.x] = (float)a;
  }
}

// Z[t, j] += r[j] on the valid samples
__global__ void add_rowvec_kernel(float* __restrict__ Z, int64_t T, int64_t ldz, int l, const float* __restrict__ r,
                                  const uint8_t* __restrict__ row_valid) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < T * l) {
    int64_t t = i / l;
    int j = (int)(i % l);
    if (!row_valid || row_valid[t]) Z[t * ldz + j] += r[j];
  }
}

// ------------------------------------------------------------------------------------------------
// project_S, SIMT: block tile 128 (s) x 64 (j), K-step 16 rows of t, 256 threads each 4 s x 8 j.
constexpr int PS_BS = 128, PS_BJ = 64, PS_BK = 16;

template <bool VEC>
__global__ void __launch_bounds__(256)
project_S_simt_kernel(const float* __restrict__ X, int64_t T, int64_t S, int64_t ldx, const float* __restrict__ pivot,
                      const float* __restrict__ dscale, const float* __restrict__ ccorr, const float* __restrict__ W,
                      int64_t ldw, int lp, const float* __restrict__ wsum, float* __restrict__ Yt, int64_t ldy) {
  __shared__ __align__(16) float Xs[PS_BK][PS_BS];
  __shared__ __align__(16) float Ws[PS_BK][PS_BJ];
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;
  const int64_t s_blk = (int64_t)blockIdx.x * PS_BS;
  const int j_blk = blockIdx.y * PS_BJ;

  // loader mapping: thread -> (row r = tid/32 (+8), 4 columns at 4*(tid%32))
  const int64_t ls = s_blk + 4 * tx;
  float pv[4];
#pragma unroll
  for (int v = 0; v < 4; ++v) pv[v] = (ls + v < S) ? pivot[ls + v] : 0.f;

  float acc[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

  for (int64_t t0 = 0; t0 < T; t0 += PS_BK) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = ty + 8 * rr;
      const int64_t t = t0 + r;
      float x[4] = {0.f, 0.f, 0.f, 0.f};
      if (t < T) {
        const float* src = X + t * ldx + ls;
        if (VEC && ls + 3 < S) {
          float4 q = ldg_stream4(src);
          x[0] = shifted(q.x, pv[0]); x[1] = shifted(q.y, pv[1]); x[2] = shifted(q.z, pv[2]); x[3] = shifted(q.w, pv[3]);
        } else {
#pragma unroll
          for (int v = 0; v < 4; ++v) if (ls + v < S) x[v] = shifted(ldg_stream1(src + v), pv[v]);
        }
      }
      *reinterpret_cast<float4*>(&Xs[r][4 * tx]) = make_float4(x[0], x[1], x[2], x[3]);
    }
    {
      // W tile: 16 rows x 64 cols = 256 float4
      const int r = tid >> 4, c4 = (tid & 15) * 4;
      const int64_t t = t0 + r;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < T && j_blk + c4 < lp) q = *reinterpret_cast<const float4*>(W + t * ldw + j_blk + c4);
      *reinterpret_cast<float4*>(&Ws[r][c4]) = q;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PS_BK; ++k) {
      const float4 xv = *reinterpret_cast<const float4*>(&Xs[k][4 * tx]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[k][8 * ty]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[k][8 * ty + 4]);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
      const float wa[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(xa[a], wa[b], acc[a][b]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const int j = j_blk + 8 * ty + b;
    if (j >= lp) continue;
    const float ws = wsum[j];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int64_t s = ls + a;
      if (s < S) {
        const float c = ccorr ? ccorr[s] : 0.f;
        Yt[(int64_t)j * ldy + s] = fmaf(dscale[s], acc[a][b], c * ws);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// project_T, SIMT: block tile 64 (t) x 64 (j) over an S range, K-step 32 columns of s, 256 threads 4 t x 4 j.
constexpr int PT_BT = 64, PT_BJ = 64, PT_BK = 32, PT_PAD = 4;

template <bool VEC>
__global__ void __launch_bounds__(256)
project_T_simt_kernel(const float* __restrict__ X, int64_t T, int64_t S, int64_t ldx, const float* __restrict__ pivot,
                      const float* __restrict__ dscale, const float* __restrict__ Yt, int64_t ldy, int lp,
                      float* __restrict__ Z, int64_t ldz, int64_t s_per_block) {
  __shared__ __align__(16) float Xs[PT_BK][PT_BT + PT_PAD];  // [s][t]
  __shared__ __align__(16) float Ys[PT_BK][PT_BJ + PT_PAD];  // [s][j]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t t_blk = (int64_t)blockIdx.x * PT_BT;
  const int j_blk = blockIdx.z * PT_BJ;
  const int64_t s_begin = (int64_t)blockIdx.y * s_per_block;
  const int64_t s_end = min(S, s_begin + s_per_block);

  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  // loader mapping: 8 threads cover one row's 32 s (one float4 each); 32 rows per sweep, 2 sweeps
  const int lc = (tid & 7) * 4, lr = tid >> 3;
  for (int64_t s0 = s_begin; s0 < s_end; s0 += PT_BK) {
    const int64_t ls = s0 + lc;
    float pv[4], dv[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const bool in = ls + v < s_end;
      pv[v] = in ? pivot[ls + v] : 0.f;
      dv[v] = in ? dscale[ls + v] : 0.f;
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = lr + 32 * rr;
      const int64_t t = t_blk + r;
      float x[4] = {0.f, 0.f, 0.f, 0.f};
      if (t < T) {
        const float* src = X + t * ldx + ls;
        if (VEC && ls + 3 < s_end) {
          float4 q = ldg_stream4(src);
          x[0] = shifted(q.x, pv[0]); x[1] = shifted(q.y, pv[1]); x[2] = shifted(q.z, pv[2]); x[3] = shifted(q.w, pv[3]);
        } else {
#pragma unroll
          for (int v = 0; v < 4; ++v) if (ls + v < s_end) x[v] = shifted(ldg_stream1(src + v), pv[v]);
        }
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) Xs[lc + v][r] = x[v];
      // Y tile: row j = j_blk + r, scaled by dscale so the inner loop is a plain product
      const int j = j_blk + r;
      float y[4] = {0.f, 0.f, 0.f, 0.f};
      if (j < lp) {
        const float* src = Yt + (int64_t)j * ldy + ls;
#pragma unroll
        for (int v = 0; v < 4; ++v) if (ls + v < s_end) y[v] = src[v] * dv[v];
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) Ys[lc + v][r] = y[v];
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < PT_BK; ++k) {
      const float4 xv = *reinterpret_cast<const float4*>(&Xs[k][4 * ty]);
      const float4 yv = *reinterpret_cast<const float4*>(&Ys[k][4 * tx]);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
      const float ya[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(xa[a], ya[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int64_t t = t_blk + 4 * ty + a;
    if (t >= T) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = j_blk + 4 * tx + b;
      if (j < lp) atomicAdd(&Z[t * ldz + j], acc[a][b]);
    }
  }
}

int launch_colsum(const float* W, int64_t T, int64_t ldw, int lp, const uint8_t* row_valid, float* out,
                  cudaStream_t stream) {
  colsum_kernel<<<(lp + 31) / 32, 256, 0, stream>>>(W, T, ldw, lp, row_valid, out);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

int launch_ccorr_dot(const float* Yt, int64_t S, int64_t ldy, const float* ccorr, int lp, float* out, cudaStream_t stream) {
  ccorr_dot_kernel<<<(unsigned)lp, 256, 0, stream>>>(Yt, S, ldy, ccorr, out);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

int project_S_simt(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot, const float* dscale,
                   const float* ccorr, const uint8_t* row_valid, const float* W, int64_t ldw, int64_t l, float* Yt, int64_t ldy,
                   float* wsum /* lp floats of workspace */, cudaStream_t stream) {
  const int lp = (int)lpad(l);
  colsum_kernel<<<(lp + 31) / 32, 256, 0, stream>>>(W, T, ldw, lp, row_valid, wsum);
  XB_LAUNCH_CHECK();
  const bool vec = (ldx % 4 == 0) && ((uintptr_t)X % 16 == 0);
  dim3 grid((unsigned)ceil_div(S, PS_BS), (unsigned)ceil_div(lp, PS_BJ));
  if (vec)
    project_S_simt_kernel<true><<<grid, 256, 0, stream>>>(X, T, S, ldx, pivot, dscale, ccorr, W, ldw, lp, wsum, Yt, ldy);
  else
    project_S_simt_kernel<false><<<grid, 256, 0, stream>>>(X, T, S, ldx, pivot, dscale, ccorr, W, ldw, lp, wsum, Yt, ldy);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

int project_T_finish(const float* Yt, int64_t T, int64_t S, int64_t ldy, const float* ccorr, const uint8_t* row_valid,
                     int64_t l, float* Z, int64_t ldz, float* rvec /* lp floats */, cudaStream_t stream) {
  if (!ccorr) return XEOFS_OK;
  ccorr_dot_kernel<<<(unsigned)l, 256, 0, stream>>>(Yt, S, ldy, ccorr, rvec);
  XB_LAUNCH_CHECK();
  add_rowvec_kernel<<<(unsigned)ceil_div(T * l, 256), 256, 0, stream>>>(Z, T, ldz, (int)l, rvec, row_valid);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

int project_T_simt(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot, const float* dscale,
                   const float* ccorr, const uint8_t* row_valid, const float* Yt, int64_t ldy, int64_t l, float* Z, int64_t ldz,
                   float* rvec, cudaStream_t stream) {
  const int lp = (int)lpad(l);
  XB_CUDA(cudaMemsetAsync(Z, 0, (size_t)T * ldz * sizeof(float), stream));
  const bool vec = (ldx % 4 == 0) && ((uintptr_t)X % 16 == 0);
  const int64_t t_tiles = ceil_div(T, PT_BT), j_tiles = ceil_div(lp, PT_BJ);
  int64_t splits = ceil_div(8 * (int64_t)num_sms(), t_tiles * j_tiles);
  int64_t spb = round_up(ceil_div(S, splits), PT_BK);
  if (spb < 4 * PT_BK) spb = 4 * PT_BK;
  splits = ceil_div(S, spb);
  XB_CHECK_ARG(splits <= 65535 && j_tiles <= 65535, "project_T: grid too large");
  dim3 grid((unsigned)t_tiles, (unsigned)splits, (unsigned)j_tiles);
  if (vec)
    project_T_simt_kernel<true><<<grid, 256, 0, stream>>>(X, T, S, ldx, pivot, dscale, Yt, ldy, lp, Z, ldz, spb);
  else
    project_T_simt_kernel<false><<<grid, 256, 0, stream>>>(X, T, S, ldx, pivot, dscale, Yt, ldy, lp, Z, ldz, spb);
  XB_LAUNCH_CHECK();
  return project_T_finish(Yt, T, S, ldy, ccorr, row_valid, l, Z, ldz, rvec, stream);
}

}  // namespace xb
