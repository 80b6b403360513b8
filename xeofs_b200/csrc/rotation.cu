// Varimax sweep kernels (R1) — see include/xeofs_b200.h.  Reference: linalg/_numpy/_rotation.py:155-177.
#include <float.h>

#include "common.cuh"

namespace xb {

// h[s] = ||L[:, s]||_2, rownorm[s] = 1 / (h + eps)   (Kaiser normalisation, _rotation.py:155-160)
__global__ void col_norms_kernel(const float* __restrict__ L, int64_t S, int m, int64_t ld, float* __restrict__ h,
                                 float* __restrict__ rownorm, float* __restrict__ Ln, int64_t ldn) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
    double a = 0.0;
    for (int j = 0; j < m; ++j) {
      const double v = (double)L[(int64_t)j * ld + s];
      a = fma(v, v, a);
    }
    const float hh = (float)sqrt(a);
    if (h) h[s] = hh;
    // the reference adds finfo(float64).eps (its loadings are fp64): only there to keep 0/0 out
    const float rn = 1.0f / (hh + 2.220446e-16f);
    if (rownorm) rownorm[s] = rn;
    if (Ln)
      for (int j = 0; j < m; ++j) Ln[(int64_t)j * ldn + s] = L[(int64_t)j * ld + s] * rn;
  }
}

// One pass over the loadings: B = Ln R (chunk of 32 features at a time), Gout += Ln^T B^3, Wout += colsum(B^2), all
// arithmetic in fp64 on the fp32 loadings: the reference's stopping rule (relative change of sum(svals) below 1e-8,
// _rotation.py:176) sits below fp32 noise — with B in fp32 the iteration stops ~15 % early and the rotated variances
// are off by 1e-3.  16x16 threads, register tile TI x TI of Gout per thread.
constexpr int VM_CHUNK = 32;

template <int TI>
__global__ void __launch_bounds__(256)
varimax_accumulate_kernel(const float* __restrict__ L, int64_t S, int m, int64_t ld, const float* __restrict__ rownorm,
                          const double* __restrict__ R, float power, const double* __restrict__ colscale,
                          double* __restrict__ Gout, double* __restrict__ Wout, float* __restrict__ absmax) {
  constexpr int MP = 16 * TI;
  extern __shared__ __align__(16) unsigned char smraw[];
  double* Rs = reinterpret_cast<double*>(smraw);          // [MP][MP+1]  R
  double* Bs = Rs + MP * (MP + 1);                         // [VM_CHUNK][MP+4]   B, then f(B)
  float* Ls = reinterpret_cast<float*>(Bs + VM_CHUNK * (MP + 4));  // [VM_CHUNK][MP+4]   normalised loadings chunk
  const int tid = threadIdx.x;
  const int ti = tid >> 4, tj = tid & 15;
  for (int idx = tid; idx < MP * MP; idx += 256) {
    const int i = idx / MP, j = idx % MP;
    Rs[i * (MP + 1) + j] = (i < m && j < m) ? R[(int64_t)i * m + j] : 0.0;
  }
  double acc[TI][TI];
#pragma unroll
  for (int a = 0; a < TI; ++a)
#pragma unroll
    for (int b = 0; b < TI; ++b) acc[a][b] = 0.0;
  double wacc[(MP + 255) / 256 > 0 ? (MP + 255) / 256 : 1];
  wacc[0] = 0.0;
  float amax = 0.f;
  const double cs = (colscale && tid < m) ? colscale[tid] : 1.0;
  (void)cs;

  const int64_t n_chunks = (S + VM_CHUNK - 1) / VM_CHUNK;
  for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const int64_t s0 = ch * VM_CHUNK;
    __syncthreads();
    {
      const int lane = tid & 31, w = tid >> 5;
      const int64_t s = s0 + lane;
      const float rn = (s < S) ? (rownorm ? rownorm[s] : 1.f) : 0.f;
      for (int j = w; j < MP; j += 8) {
        float v = 0.f;
        if (j < m && s < S) v = L[(int64_t)j * ld + s] * rn;
        Ls[lane * (MP + 4) + j] = v;
      }
    }
    __syncthreads();
    // B[r][j] = sum_i Ls[r][i] Rs[i][j];  thread -> (r = tid / 8 -> 32 rows, 8 threads per row, each MP/8 columns)
    {
      const int r = tid >> 3, c0 = tid & 7;
      for (int j = c0; j < MP; j += 8) {
        double b = 0.0;
        for (int i = 0; i < m; ++i) b = fma((double)Ls[r * (MP + 4) + i], Rs[i * (MP + 1) + j], b);
        Bs[r * (MP + 4) + j] = b;
      }
    }
    __syncthreads();
    // column sums of B^2 (fp64) by the first MP threads, then cube in place
    if (tid < MP) {
      double w2 = 0.0;
#pragma unroll 4
      for (int r = 0; r < VM_CHUNK; ++r) {
        const double b = Bs[r * (MP + 4) + tid];
        w2 = fma(b, b, w2);
        amax = fmaxf(amax, fabsf((float)b));
        // f(b): varimax b^3; promax target (b c)|b c|^(power-1)   (_rotation.py:57-62, 166-170)
        double fb;
        if (power == 3.f && !colscale) fb = b * b * b;
        else { const double u = b * cs; fb = (power == 1.f) ? u : u * pow(fabs(u), (double)power - 1.0); }
        Bs[r * (MP + 4) + tid] = fb;
      }
      wacc[0] += w2;
    }
    __syncthreads();
#pragma unroll 2
    for (int r = 0; r < VM_CHUNK; ++r) {
      double a[TI], b[TI];
#pragma unroll
      for (int x = 0; x < TI; ++x) {
        a[x] = (double)Ls[r * (MP + 4) + ti * TI + x];
        b[x] = Bs[r * (MP + 4) + tj * TI + x];
      }
#pragma unroll
      for (int x = 0; x < TI; ++x)
#pragma unroll
        for (int y = 0; y < TI; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
    }
  }
#pragma unroll
  for (int x = 0; x < TI; ++x)
#pragma unroll
    for (int y = 0; y < TI; ++y) {
      const int i = ti * TI + x, j = tj * TI + y;
      if (i < m && j < m) atomicAdd(&Gout[(int64_t)i * m + j], acc[x][y]);
    }
  if (tid < m) {
    atomicAdd(&Wout[tid], wacc[0]);
    if (absmax) atomicMax(reinterpret_cast<int*>(&absmax[tid]), __float_as_int(amax));  // non-negative floats
  }
}

// The varimax case of the same pass (power 3, no column scale) on the fp64 tensor-core path: both products as
// mma.sync m8n8k4 (IEEE fp64 FMAs, so the arithmetic is the one of the kernel above), 64 features per turn.
//   GEMM1  B[s, j] = sum_i L[s, i] R[i, j]      warp w owns the 8 features 8w.., all MT column tiles
//   f = b^3 (in the accumulator registers), W[j] += b^2 (shuffle over the 8 rows, shared-memory atomics)
//   GEMM2  G[i, j] += sum_s L[s, i] f[s, j]     warp w owns the row tiles w and w + 8, accumulators live in registers
//                                               over all the turns of the CTA
// Shared memory: R, the chunk of loadings and f, all fp64, rows padded to a pitch of 4 mod 16 doubles (the fragment
// loads of a half-warp then fall into 16 different bank pairs).
constexpr int VX_CH = 64;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int MT>  // 8-wide tiles of the mode axis (m <= 8 MT)
__global__ void __launch_bounds__(256, 1)
varimax_exact_mma_kernel(const float* __restrict__ L, int64_t S, int m, int64_t ld, const float* __restrict__ rownorm,
                         const double* __restrict__ R, double* __restrict__ Gout, double* __restrict__ Wout) {
  constexpr int MP = 8 * MT, LDP = MP + 4, MW = (MT + 7) / 8;
  extern __shared__ __align__(16) unsigned char smraw[];
  double* Rs = reinterpret_cast<double*>(smraw);  // [MP][LDP]
  double* Ls = Rs + MP * LDP;                     // [VX_CH][LDP]
  double* Fs = Ls + VX_CH * LDP;                  // [VX_CH][LDP]
  double* Wsh = Fs + VX_CH * LDP;                 // [MP]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lc = lane & 3;
  for (int idx = tid; idx < MP * MP; idx += 256) {
    const int i = idx / MP, j = idx % MP;
    Rs[i * LDP + j] = (i < m && j < m) ? R[(int64_t)i * m + j] : 0.0;
  }
  if (tid < MP) Wsh[tid] = 0.0;
  double acc2[MW][MT][2];
#pragma unroll
  for (int a = 0; a < MW; ++a)
#pragma unroll
    for (int b = 0; b < MT; ++b) acc2[a][b][0] = acc2[a][b][1] = 0.0;
  const int ksteps = (m + 3) >> 2;

  const int64_t n_chunks = (S + VX_CH - 1) / VX_CH;
  for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const int64_t s0 = ch * VX_CH;
    __syncthreads();  // GEMM2 of the previous turn has read Ls and Fs (first turn: Rs is written)
    {
      const int sl = tid & 63, w4 = tid >> 6;
      const int64_t s = s0 + sl;
      const float rn = (s < S) ? (rownorm ? rownorm[s] : 1.f) : 0.f;
      for (int j = w4; j < MP; j += 4) {
        float v = 0.f;
        if (j < m && s < S) v = L[(int64_t)j * ld + s] * rn;
        Ls[sl * LDP + j] = (double)v;
      }
    }
    __syncthreads();
    {
      double acc1[MT][2];
#pragma unroll
      for (int b = 0; b < MT; ++b) acc1[b][0] = acc1[b][1] = 0.0;
      const double* arow = Ls + (8 * warp + lr) * LDP + lc;
      const double* bcol = Rs + lc * LDP + lr;
      for (int k = 0; k < ksteps; ++k) {
        const double a = arow[4 * k];
#pragma unroll
        for (int b = 0; b < MT; ++b) dmma884(acc1[b][0], acc1[b][1], a, bcol[4 * k * LDP + 8 * b]);
      }
      // accumulator element e of tile b: feature 8 warp + lr, mode 8 b + 2 lc + e
#pragma unroll
      for (int b = 0; b < MT; ++b) {
        const double b0 = acc1[b][0], b1 = acc1[b][1];
        double w0 = b0 * b0, w1 = b1 * b1;
        *reinterpret_cast<double2*>(Fs + (8 * warp + lr) * LDP + 8 * b + 2 * lc) = make_double2(w0 * b0, w1 * b1);
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          w0 += __shfl_xor_sync(0xffffffffu, w0, o);
          w1 += __shfl_xor_sync(0xffffffffu, w1, o);
        }
        if (lr == 0) {
          atomicAdd(&Wsh[8 * b + 2 * lc], w0);
          atomicAdd(&Wsh[8 * b + 2 * lc + 1], w1);
        }
      }
    }
    __syncthreads();
#pragma unroll 2
    for (int k = 0; k < VX_CH / 4; ++k) {
      double a[MW];
#pragma unroll
      for (int x = 0; x < MW; ++x) {
        const int mt = warp + 8 * x;
        a[x] = mt < MT ? Ls[(4 * k + lc) * LDP + 8 * mt + lr] : 0.0;
      }
      const double* frow = Fs + (4 * k + lc) * LDP + lr;
#pragma unroll
      for (int b = 0; b < MT; ++b) {
        const double fb = frow[8 * b];
#pragma unroll
        for (int x = 0; x < MW; ++x)
          if (warp + 8 * x < MT) dmma884(acc2[x][b][0], acc2[x][b][1], a[x], fb);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int x = 0; x < MW; ++x) {
    const int mt = warp + 8 * x;
    if (mt >= MT) continue;
    const int i = 8 * mt + lr;
#pragma unroll
    for (int b = 0; b < MT; ++b) {
      const int j = 8 * b + 2 * lc;
      if (i < m && j < m) atomicAdd(&Gout[(int64_t)i * m + j], acc2[x][b][0]);
      if (i < m && j + 1 < m) atomicAdd(&Gout[(int64_t)i * m + j + 1], acc2[x][b][1]);
    }
  }
  if (tid < m) atomicAdd(&Wout[tid], Wsh[tid]);
}

template <int MT>
static int launch_exact_mma(const float* L, int64_t S, int m, int64_t ld, const float* rownorm, const double* R,
                            double* Gout, double* Wout, cudaStream_t stream) {
  constexpr int MP = 8 * MT, LDP = MP + 4;
  const size_t smem = ((size_t)MP * LDP + 2 * (size_t)VX_CH * LDP + MP) * sizeof(double);
  XB_CUDA(cudaFuncSetAttribute(varimax_exact_mma_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks = (int)imin(ceil_div(S, VX_CH), (int64_t)num_sms());
  varimax_exact_mma_kernel<MT><<<blocks, 256, smem, stream>>>(L, S, m, ld, rownorm, R, Gout, Wout);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

// project_tc.cu
int env_int(const char* name, int dflt);
// rotation_tc.cu
int64_t varimax_tc_workspace_bytes(int64_t S, int64_t m);
bool varimax_tc_supported(const float* L, int64_t S, int64_t m, int64_t ld);
int varimax_sweep_tc(const float* L, const float* packed, int64_t S, int64_t m, int64_t ld, const double* R, double* Gout,
                     double* Wout, int accumulate, int products, void* workspace, int64_t workspace_bytes,
                     cudaStream_t stream);
int64_t varimax_pack_bytes(int64_t S, int64_t m);
int varimax_pack(const float* L, int64_t S, int64_t m, int64_t ld, float* packed, cudaStream_t stream);

}  // namespace xb

using namespace xb;

extern "C" int64_t xeofs_b200_varimax_workspace_bytes(int64_t S, int64_t m) { return varimax_tc_workspace_bytes(S, m); }

extern "C" int64_t xeofs_b200_varimax_pack_bytes(int64_t S, int64_t m) {
  return xeofs_b200_has_tcgen05() ? varimax_pack_bytes(S, m) : 0;
}

extern "C" int xeofs_b200_varimax_pack(const float* Ln, int64_t S, int64_t m, int64_t ld, float* packed,
                                       int64_t packed_bytes, void* stream_) {
  XB_CHECK_ARG(Ln && packed && S > 0 && ld >= S && m >= 2 && m <= 128, "varimax_pack: bad arguments");
  XB_CHECK_ARG(ld % 4 == 0 && (uintptr_t)Ln % 16 == 0 && (uintptr_t)packed % 128 == 0, "varimax_pack: misaligned pointers");
  const int64_t need = xeofs_b200_varimax_pack_bytes(S, m);
  XB_CHECK_ARG(need > 0 && packed_bytes >= need, "varimax_pack: buffer too small (%lld < %lld bytes) or unsupported shape",
               (long long)packed_bytes, (long long)need);
  return varimax_pack(Ln, S, m, ld, packed, (cudaStream_t)stream_);
}

extern "C" int xeofs_b200_varimax_sweep(const float* Ln, const float* packed, int64_t S, int64_t m, int64_t ld,
                                        const double* R, double* Gout, double* Wout, int accumulate, int products,
                                        void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(Ln && R && Gout && Wout && workspace && S > 0 && ld >= S, "varimax_sweep: bad arguments");
  XB_CHECK_ARG(products == 1 || products == 3, "varimax_sweep: products must be 1 (single TF32) or 3 (3xTF32)");
  XB_CHECK_ARG(m >= 2 && m <= 128, "varimax_sweep: m=%lld must be in 2..128", (long long)m);
  XB_CHECK_ARG((uintptr_t)workspace % 256 == 0, "varimax_sweep: misaligned workspace");
  if (!xeofs_b200_has_tcgen05() || !varimax_tc_supported(Ln, S, m, ld)) {
    set_error("varimax_sweep: needs the tcgen05 path (sm_100, 16-byte aligned Ln, ld %% 4 == 0)");
    return XEOFS_E_UNSUPPORTED;
  }
  XB_CHECK_ARG(!packed || (uintptr_t)packed % 128 == 0, "varimax_sweep: misaligned packed copy");
  return varimax_sweep_tc(Ln, packed, S, m, ld, R, Gout, Wout, accumulate, products, workspace, workspace_bytes, stream);
}

extern "C" int xeofs_b200_col_norms(const float* L, int64_t S, int64_t m, int64_t ld, float* h, float* rownorm,
                                    float* Ln, int64_t ldn, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(L && S > 0 && m > 0 && ld >= S, "col_norms: bad arguments");
  const int blocks = (int)imin(ceil_div(S, 256), 8 * (int64_t)num_sms());
  XB_CHECK_ARG(!Ln || ldn >= S, "col_norms: ldn < S");
  col_norms_kernel<<<blocks, 256, 0, stream>>>(L, S, (int)m, ld, h, rownorm, Ln, ldn);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

extern "C" int xeofs_b200_varimax_accumulate(const float* L, int64_t S, int64_t m, int64_t ld, const float* rownorm,
                                             const double* R, double power, const double* colscale, double* Gout,
                                             double* Wout, float* absmax, int accumulate, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(L && R && Gout && Wout && S > 0 && ld >= S, "varimax_accumulate: bad arguments");
  XB_CHECK_ARG(m >= 2 && m <= 128, "varimax_accumulate: m=%lld must be in 2..128", (long long)m);
  if (!accumulate) {
    XB_CUDA(cudaMemsetAsync(Gout, 0, (size_t)m * m * sizeof(double), stream));
    XB_CUDA(cudaMemsetAsync(Wout, 0, (size_t)m * sizeof(double), stream));
    if (absmax) XB_CUDA(cudaMemsetAsync(absmax, 0, (size_t)m * sizeof(float), stream));
  }
  XB_CHECK_ARG(power >= 1.0, "varimax_accumulate: power must be >= 1");
  if (power == 3.0 && !colscale && !absmax && m > 32 && m <= 104 && env_int("XEOFS_VX_MMA", 1))
    return m <= 64 ? launch_exact_mma<8>(L, S, (int)m, ld, rownorm, R, Gout, Wout, stream)
                   : launch_exact_mma<13>(L, S, (int)m, ld, rownorm, R, Gout, Wout, stream);
  const int ti = m <= 16 ? 1 : m <= 32 ? 2 : m <= 64 ? 4 : 8;
  const int MP = 16 * ti;
  const size_t smem = ((size_t)MP * (MP + 1) + (size_t)VM_CHUNK * (MP + 4)) * sizeof(double) +
                      (size_t)VM_CHUNK * (MP + 4) * sizeof(float);
  const int blocks = (int)imin(ceil_div(S, VM_CHUNK), 2 * (int64_t)num_sms());
#define XB_VM(TI)                                                                                                  \
  XB_CUDA(cudaFuncSetAttribute(varimax_accumulate_kernel<TI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
  varimax_accumulate_kernel<TI><<<blocks, 256, smem, stream>>>(L, S, (int)m, ld, rownorm, R, (float)power, colscale, Gout, Wout, absmax)
  switch (ti) {
    case 1: XB_VM(1); break;
    case 2: XB_VM(2); break;
    case 4: XB_VM(4); break;
    default: XB_VM(8); break;
  }
#undef XB_VM
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}
