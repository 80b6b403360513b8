// inverse_transform: X_rec[t,s] = (sum_m scores[t,m] * Vt[modes[m], s] - ccorr[s]) / dscale[s] + pivot[s], NaN at dropped
// features.  Reference: single/eof.py:134-156 (components . scores), preprocessing/sanitizer.py:128-153 (reindex ->
// NaN), preprocessing/scaler.py:165-190 (/weights /coslat *std +mean) — the last three folded into the epilogue, so the
// T x S output is written exactly once.
#include "common.cuh"

namespace xb {

constexpr int RC_BT = 64, RC_BS = 128, RC_BK = 16;

__global__ void __launch_bounds__(256)
reconstruct_kernel(const float* __restrict__ Sc, int64_t T, int64_t lds, const float* __restrict__ Vt, int64_t S, int64_t ldv,
                   const int32_t* __restrict__ modes, int m, const float* __restrict__ pivot,
                   const float* __restrict__ dscale, const float* __restrict__ ccorr, const uint8_t* __restrict__ valid,
                   float* __restrict__ Out, int64_t ldo) {
  __shared__ __align__(16) float Ss[RC_BK][RC_BT + 4];  // [mode][t]
  __shared__ __align__(16) float Vs[RC_BK][RC_BS];      // [mode][s]
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;  // thread tile: 8 t (ty*8..) x 4 s (tx*4..)
  const int64_t s_blk = (int64_t)blockIdx.x * RC_BS, t_blk = (int64_t)blockIdx.y * RC_BT;
  float acc[8][4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int m0 = 0; m0 < m; m0 += RC_BK) {
    // scores tile: 64 t x 16 modes
    for (int idx = tid; idx < RC_BT * RC_BK; idx += 256) {
      const int r = idx / RC_BK, c = idx % RC_BK;
      const int64_t t = t_blk + r;
      Ss[c][r] = (t < T && m0 + c < m) ? Sc[t * lds + m0 + c] : 0.f;
    }
    // component tile: 16 modes x 128 s
    for (int idx = tid; idx < RC_BK * RC_BS; idx += 256) {
      const int r = idx / RC_BS, c = idx % RC_BS;
      const int64_t s = s_blk + c;
      Vs[r][c] = (m0 + r < m && s < S) ? Vt[(int64_t)modes[m0 + r] * ldv + s] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RC_BK; ++k) {
      const float4 v = *reinterpret_cast<const float4*>(&Vs[k][4 * tx]);
      const float4 s0 = *reinterpret_cast<const float4*>(&Ss[k][8 * ty]);
      const float4 s1 = *reinterpret_cast<const float4*>(&Ss[k][8 * ty + 4]);
      const float sa[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
      const float va[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(sa[a], va[b], acc[a][b]);
    }
    __syncthreads();
  }
  float pv[4], inv[4], cc[4];
  bool ok[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int64_t s = s_blk + 4 * tx + b;
    const bool in = s < S;
    ok[b] = in && (valid ? valid[s] != 0 : true);
    pv[b] = in ? pivot[s] : 0.f;
    cc[b] = (in && ccorr) ? ccorr[s] : 0.f;
    inv[b] = ok[b] ? 1.0f / dscale[s] : 0.f;
  }
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int64_t t = t_blk + 8 * ty + a;
    if (t >= T) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t s = s_blk + 4 * tx + b;
      if (s < S) Out[t * ldo + s] = ok[b] ? fmaf(acc[a][b] - cc[b], inv[b], pv[b]) : nanf("");
    }
  }
}

}  // namespace xb

using namespace xb;

extern "C" int xeofs_b200_reconstruct(const float* scores, int64_t T, int64_t lds, const float* Vt, int64_t S, int64_t ldv,
                                      const int32_t* modes, int64_t m, const float* pivot, const float* dscale,
                                      const float* ccorr, const uint8_t* valid, float* out, int64_t ldo, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(scores && Vt && modes && pivot && dscale && out, "reconstruct: null pointer");
  XB_CHECK_ARG(T > 0 && S > 0 && m > 0 && lds >= m && ldv >= S && ldo >= S, "reconstruct: bad shape");
  const int64_t gx = ceil_div(S, RC_BS), gy = ceil_div(T, RC_BT);
  XB_CHECK_ARG(gy <= 65535, "reconstruct: too many samples for one launch (%lld)", (long long)T);
  reconstruct_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, stream>>>(scores, T, lds, Vt, S, ldv, modes, (int)m, pivot,
                                                                          dscale, ccorr, valid, out, ldo);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

namespace xb {
// out[j, s] = A[t0 + j, s] for j < nrows (the preprocessed samples as a space-side block), zero for the pad rows
__global__ void scaled_rows_kernel(const float* __restrict__ X, int64_t S, int64_t ldx, const float* __restrict__ pivot,
                                   const float* __restrict__ dscale, const float* __restrict__ ccorr,
                                   const uint8_t* __restrict__ row_valid, int64_t t0, int nrows, float* __restrict__ out,
                                   int64_t ldo) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (s >= S) return;
  float v = 0.f;
  if (j < nrows && (!row_valid || row_valid[t0 + j])) {
    const float d = X[(t0 + j) * ldx + s] - pivot[s];
    v = ((d == d) ? d * dscale[s] : 0.f) + (ccorr ? ccorr[s] : 0.f);
  }
  out[(int64_t)j * ldo + s] = v;
}
}  // namespace xb

namespace xb {
__global__ void __launch_bounds__(256)
materialize_kernel(const float* __restrict__ X, int64_t T, int64_t S, int64_t ldx, const float* __restrict__ pivot,
                   const float* __restrict__ dscale, const float* __restrict__ ccorr, const uint8_t* __restrict__ row_valid,
                   int64_t rows_out, int round_tf32, float* __restrict__ out, int64_t ldo) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const float p = pivot[s], d = dscale[s], c = ccorr ? ccorr[s] : 0.f;
  for (int64_t t = blockIdx.y; t < rows_out; t += gridDim.y) {
    float v = 0.f;
    if (t < T && (!row_valid || row_valid[t])) {
      const float x = X[t * ldx + s] - p;
      v = ((x == x) ? x * d : 0.f) + c;
      if (round_tf32) {
        const uint32_t u = __float_as_uint(v);
        v = __uint_as_float((u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u);
      }
    }
    out[t * ldo + s] = v;
  }
}
}  // namespace xb

extern "C" int xeofs_b200_materialize(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot,
                                      const float* dscale, const float* ccorr, const uint8_t* row_valid, int64_t rows_out,
                                      int round_tf32, float* out, int64_t ldo, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(X && pivot && dscale && out, "materialize: null pointer");
  XB_CHECK_ARG(T > 0 && S > 0 && ldx >= S && ldo >= S && rows_out >= T, "materialize: bad shape");
  materialize_kernel<<<dim3((unsigned)ceil_div(S, 256), (unsigned)imin(rows_out, 4096)), 256, 0, stream>>>(
      X, T, S, ldx, pivot, dscale, ccorr, row_valid, rows_out, round_tf32, out, ldo);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

extern "C" int xeofs_b200_scaled_rows(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot,
                                      const float* dscale, const float* ccorr, const uint8_t* row_valid, int64_t t0,
                                      int64_t nrows, int64_t rows_out, float* out, int64_t ldo, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(X && pivot && dscale && out, "scaled_rows: null pointer");
  XB_CHECK_ARG(S > 0 && ldx >= S && ldo >= S && t0 >= 0 && nrows > 0 && t0 + nrows <= T && rows_out >= nrows &&
                   rows_out <= 65535,
               "scaled_rows: bad shape");
  scaled_rows_kernel<<<dim3((unsigned)ceil_div(S, 256), (unsigned)rows_out), 256, 0, stream>>>(
      X, S, ldx, pivot, dscale, ccorr, row_valid, t0, (int)nrows, out, ldo);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}
