// Column statistics (one streaming read of X) and the Scaler vectors derived from them.
// Reference spans replaced: preprocessing/scaler.py:100-116, preprocessing/sanitizer.py:46-56,
// utils/xarray_utils.py:236-253 (see include/xeofs_b200.h).
#include <float.h>

#include "common.cuh"

namespace xb {

constexpr int kStatsThreads = 256;
constexpr int kStatsWarps = kStatsThreads / 32;

// Block = 8 warps over a (rows_per_block x 32*VEC) slab; warp w takes rows w, w+8, ...; lane owns VEC
// adjacent columns, so every warp-wide load is one contiguous 128*VEC-byte run of a row.
template <int VEC>
__global__ void __launch_bounds__(kStatsThreads)
col_stats_kernel(const float* __restrict__ X, int64_t T, int64_t S, int64_t ldx, int64_t rows_per_block,
                 float* __restrict__ shift, double* __restrict__ sum, double* __restrict__ sumsq,
                 int32_t* __restrict__ cnt, int32_t* __restrict__ row_nan) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t s0 = ((int64_t)blockIdx.x * 32 + lane) * VEC;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(T, r0 + rows_per_block);

  float p[VEC], a1[VEC], a2[VEC];
  int n[VEC];
  bool inb[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    inb[v] = (s0 + v) < S;
    float x0 = inb[v] ? X[s0 + v] : 0.f;  // row 0 as the shift keeps the shifted sums small
    p[v] = (x0 == x0) ? x0 : 0.f;
    a1[v] = 0.f; a2[v] = 0.f; n[v] = 0;
  }
  if (blockIdx.y == 0 && warp == 0) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) if (inb[v]) shift[s0 + v] = p[v];
  }

  constexpr int UNR = 4;  // independent row loads in flight per thread
  for (int64_t tb = r0 + warp; tb < r1; tb += kStatsWarps * UNR) {
    float x[UNR][VEC];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t t = tb + (int64_t)u * kStatsWarps;
      const float* row = X + t * ldx + s0;
      const bool rok = t < r1;
      if constexpr (VEC == 4) {
        if (rok && inb[3]) {
          float4 q = ldg_stream4(row);
          x[u][0] = q.x; x[u][1] = q.y; x[u][2] = q.z; x[u][3] = q.w;
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) x[u][v] = (rok && inb[v]) ? ldg_stream1(row + v) : p[v];
        }
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) x[u][v] = (rok && inb[v]) ? ldg_stream1(row + v) : p[v];
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t t = tb + (int64_t)u * kStatsWarps;
      const bool rok = t < r1;
      int nan_here = 0;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float d = x[u][v] - p[v];
        bool ok = (d == d);
        const bool live = rok && inb[v];
        nan_here += (live && !ok) ? 1 : 0;
        d = (ok && live) ? d : 0.f;
        a1[v] += d;
        a2[v] = fmaf(d, d, a2[v]);
        n[v] += (ok && live) ? 1 : 0;
      }
      if (__any_sync(0xffffffffu, nan_here)) {
        int tot = warp_sum(nan_here);
        if (lane == 0) atomicAdd(&row_nan[t], tot);
      }
    }
  }

  __shared__ double s1[kStatsWarps][32 * VEC];
  __shared__ double s2[kStatsWarps][32 * VEC];
  __shared__ int sn[kStatsWarps][32 * VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    s1[warp][lane * VEC + v] = (double)a1[v];
    s2[warp][lane * VEC + v] = (double)a2[v];
    sn[warp][lane * VEC + v] = n[v];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * VEC; c += kStatsThreads) {
    int64_t s = (int64_t)blockIdx.x * 32 * VEC + c;
    if (s >= S) continue;
    double t1 = 0, t2 = 0;
    int tn = 0;
#pragma unroll
    for (int w = 0; w < kStatsWarps; ++w) { t1 += s1[w][c]; t2 += s2[w][c]; tn += sn[w][c]; }
    atomicAdd(&sum[s], t1);
    atomicAdd(&sumsq[s], t2);
    atomicAdd(&cnt[s], tn);
  }
}

// NOTE: a block that does not own row 0 still shifts by X[0, s]: every block reads row 0 itself.
__global__ void scaling_finalize_kernel(int64_t S, const float* __restrict__ shift, const double* __restrict__ sum,
                                        const double* __restrict__ sumsq, const int32_t* __restrict__ cnt,
                                        const double* __restrict__ featw, int flags, float* __restrict__ mean,
                                        float* __restrict__ stdv, uint8_t* __restrict__ valid,
                                        float* __restrict__ pivot, float* __restrict__ dscale,
                                        float* __restrict__ ccorr, double* __restrict__ scalars) {
  const bool center = flags & XEOFS_F_CENTER, standardize = flags & XEOFS_F_STANDARDIZE;
  double tv = 0.0;
  int nvalid = 0, cmax = 0, cmin = 0x7fffffff;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
    const int n = cnt[s];
    const bool ok = n > 0;
    double mu = 0.0, m2 = 0.0;
    if (ok) {
      const double a = sum[s] / n;
      mu = (double)shift[s] + a;
      m2 = sumsq[s] - sum[s] * a;  // sum (x - mean)^2
      if (m2 < 0) m2 = 0;
    }
    const float mu32 = (float)mu;
    // std: ddof = 0, clipped at float32 eps (scaler.py:105-108)
    float sd = ok ? fmaxf((float)sqrt(m2 / n), FLT_EPSILON) : nanf("");
    double d = ok ? (featw ? featw[s] : 1.0) : 0.0;
    if (standardize && ok) d /= (double)sd;
    const float d32 = (float)d;
    const float piv = ok ? (center ? mu32 : shift[s]) : 0.f;
    const float mu_eff = center ? mu32 : 0.f;
    if (mean) mean[s] = ok ? mu32 : nanf("");
    if (stdv) stdv[s] = sd;
    valid[s] = ok ? 1 : 0;
    pivot[s] = piv;
    dscale[s] = d32;
    ccorr[s] = ok ? (piv - mu_eff) * d32 : 0.f;
    if (ok) {
      // var(ddof=1) of the scaled column: the variance ignores any constant offset
      if (n > 1) tv += (double)d32 * (double)d32 * m2 / (double)(n - 1);
      nvalid += 1;
      cmax = max(cmax, n);
      cmin = min(cmin, n);
    }
  }
  tv = warp_sum(tv);
  nvalid = warp_sum(nvalid);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&scalars[0], tv);
    atomicAdd(&scalars[1], (double)nvalid);
    // max / min of non-negative doubles through their bit patterns
    atomicMax((unsigned long long*)&scalars[2], (unsigned long long)__double_as_longlong((double)cmax));
    atomicMin((unsigned long long*)&scalars[3], (unsigned long long)__double_as_longlong((double)cmin));
  }
}

__global__ void init_scalars_kernel(double* s) {
  if (threadIdx.x == 0) { s[0] = 0.0; s[1] = 0.0; s[2] = 0.0; s[3] = 2147483647.0; }
}

}  // namespace xb

using namespace xb;

extern "C" int xeofs_b200_col_stats(const float* X, int64_t T, int64_t S, int64_t ldx, float* shift, double* sum,
                                    double* sumsq, int32_t* cnt, int32_t* row_nan, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(X && shift && sum && sumsq && cnt && row_nan, "col_stats: null pointer");
  XB_CHECK_ARG(T > 0 && S > 0 && ldx >= S, "col_stats: bad shape T=%lld S=%lld ldx=%lld", (long long)T, (long long)S, (long long)ldx);
  XB_CUDA(cudaMemsetAsync(sum, 0, S * sizeof(double), stream));
  XB_CUDA(cudaMemsetAsync(sumsq, 0, S * sizeof(double), stream));
  XB_CUDA(cudaMemsetAsync(cnt, 0, S * sizeof(int32_t), stream));
  XB_CUDA(cudaMemsetAsync(row_nan, 0, T * sizeof(int32_t), stream));
  const bool vec = (ldx % 4 == 0) && ((uintptr_t)X % 16 == 0);
  const int cols = vec ? 128 : 32;
  const int64_t col_blocks = ceil_div(S, cols);
  // enough row splits to give every SM a few blocks, but long enough runs per thread
  int64_t want = 4 * (int64_t)num_sms();
  int64_t row_splits = col_blocks >= want ? 1 : ceil_div(want, col_blocks);
  int64_t rpb = round_up(ceil_div(T, row_splits), kStatsWarps);
  if (rpb < 64) rpb = 64;
  row_splits = ceil_div(T, rpb);
  XB_CHECK_ARG(row_splits <= 65535, "col_stats: too many row splits");
  dim3 grid((unsigned)col_blocks, (unsigned)row_splits);
  if (vec)
    col_stats_kernel<4><<<grid, kStatsThreads, 0, stream>>>(X, T, S, ldx, rpb, shift, sum, sumsq, cnt, row_nan);
  else
    col_stats_kernel<1><<<grid, kStatsThreads, 0, stream>>>(X, T, S, ldx, rpb, shift, sum, sumsq, cnt, row_nan);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

extern "C" int xeofs_b200_scaling_finalize(int64_t S, const float* shift, const double* sum, const double* sumsq,
                                           const int32_t* cnt, const double* featw, int flags, float* mean,
                                           float* stdv, uint8_t* valid, float* pivot, float* dscale, float* ccorr,
                                           double* scalars_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(S > 0 && shift && sum && sumsq && cnt && valid && pivot && dscale && ccorr && scalars_out,
               "scaling_finalize: null pointer");
  init_scalars_kernel<<<1, 32, 0, stream>>>(scalars_out);
  int blocks = (int)imin(ceil_div(S, 256), 8 * (int64_t)num_sms());
  scaling_finalize_kernel<<<blocks, 256, 0, stream>>>(S, shift, sum, sumsq, cnt, featw, flags, mean, stdv, valid,
                                                      pivot, dscale, ccorr, scalars_out);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}
