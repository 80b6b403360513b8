// Sample Gram matrix A A^T of a preprocessed field on the tensor cores, for the total squared covariance of the cross
// models:  sum |C|^2 = sum |X^T Y|^2 / (n-1)^2 = <X X^T, Y Y^T>_F / (n-1)^2   (reference: cross/cpcca.py:991-1000 forms
// C = X^T Y / (n-1) densely, cpcca.py:1008-1015 — at BASELINE config 3 that matrix has 6.7e10 entries).
//
// A plain tcgen05 GEMM: both operands are row tiles of ONE bf16 copy of the preprocessed matrix (xeofs_b200_materialize_bf16:
// T_pad x S_pad, rows = samples, features contiguous), fetched by TMA with the 128-byte swizzle straight into the K-major
// layout kind::f16 reads — no register staging.  Every entry of the Gram matrix is a sum over >= 65 536 features of
// products whose bf16 rounding errors (unit roundoff 2^-8, round-to-nearest) are independent: they average out to ~1e-5
// relative per entry and far below that in the sum over the T^2 entries that is read from them.
//
//   D[128 t][256 t'] += A[t, s-chunk] . A[t', s-chunk]^T       M = 128, N = 256, K = 16 per instruction, 64 per stage
//
// Grid per panel of 256 columns t': (row tiles with t >= panel start) x (splits of the feature axis, for load balance);
// the partial sums of the splits are added by a second, deterministic kernel.  Only the block-lower triangle is
// computed (the matrix is symmetric).
//
// The tensor core adds into its fp32 accumulator with truncation (measured here: -6.8e-8 of the sum per instruction,
// -3.4e-5 after 508): two TMEM accumulators take turns, each moved after GB_FLUSH stages (64 instructions) into fp32
// registers of the epilogue warps, whose adds round to nearest.
//
// 10 warps: warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue (TMEM -> registers -> global; two warps per
// TMEM lane quarter, 128 columns each).
#include <cuda.h>
#include <cuda_bf16.h>

#include "tc_common.cuh"

namespace xb {

constexpr int GB_KS = 64;          // features per stage (= one 128-byte swizzle row of bf16)
constexpr int GB_N = 256;          // columns of the panel
constexpr int GB_STAGES = 4;
constexpr int GB_A_BYTES = TC_TILE * GB_KS * 2;  // 16 KB
constexpr int GB_B_BYTES = GB_N * GB_KS * 2;     // 32 KB
constexpr int GB_FLUSH = 16;       // stages (4 instructions each) a TMEM accumulator sums before it moves to registers

struct GbParams {
  int row_tile0;   // first row tile (of 128 rows) of this launch
  int col0;        // first column t' of the panel
  int ksteps;      // stages' worth of K over the whole feature axis
  int ksteps_per_split;
  float* part;     // [split][n_row_tiles * 128][256]
  int64_t rows;    // n_row_tiles * 128
};

__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// D fp32, A / B bf16, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_TILE >> 4) << 24);
}

__global__ void __launch_bounds__(320, 1)
gram_bf16_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GbParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* as = smem;                                   // [stages][16 KB]
  uint8_t* bs = smem + GB_STAGES * GB_A_BYTES;          // [stages][32 KB]
  uint64_t* bars = (uint64_t*)(bs + GB_STAGES * GB_B_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + GB_STAGES;
  uint64_t* dfull = bars + 2 * GB_STAGES;   // [2] accumulator buffer complete
  uint64_t* dempty = dfull + 2;             // [2] accumulator buffer moved into registers
  uint32_t* tmem_slot = (uint32_t*)(dempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (p.row_tile0 + (int)blockIdx.x) * TC_TILE;
  const int k0 = (int)blockIdx.y * p.ksteps_per_split;
  const int nk = min(p.ksteps_per_split, p.ksteps - k0);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < GB_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&dfull[b], 1);
      mbar_init(&dempty[b], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * GB_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    Pipe pp;
    for (int c = 0; c < nk; ++c, pp.advance(GB_STAGES)) {
      mbar_wait(&empty[pp.st], pp.ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full[pp.st], GB_A_BYTES + GB_B_BYTES);
        const int s0 = (k0 + c) * GB_KS;
        tma_load_2d(as + pp.st * GB_A_BYTES, &mapA, s0, row0, &full[pp.st], HINT_EVICT_FIRST);
        tma_load_2d(bs + pp.st * GB_B_BYTES, &mapB, s0, p.col0, &full[pp.st], HINT_EVICT_LAST);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(GB_N);
    Pipe pp;
    int g = 0, cg = 0;  // flush group and stage within it
    for (int c = 0; c < nk; ++c, pp.advance(GB_STAGES)) {
      const int buf = g & 1;
      if (cg == 0) mbar_wait(&dempty[buf], (((uint32_t)g >> 1) & 1) ^ 1);  // the registers hold what this buffer had
      mbar_wait(&full[pp.st], pp.ph);
      tc_fence_after();
      const bool group_end = cg + 1 == GB_FLUSH || c == nk - 1;
      if (elect_one()) {
        const uint64_t da = make_b_desc(smem_u32(as + pp.st * GB_A_BYTES));
        const uint64_t db = make_b_desc(smem_u32(bs + pp.st * GB_B_BYTES));
#pragma unroll
        for (int k = 0; k < GB_KS / 16; ++k)  // +32 bytes (16 bf16) along K = +2 in the (address >> 4) field
          mma_f16_ss(tmem_base + buf * GB_N, da + 2 * k, db + 2 * k, idesc, !(cg == 0 && k == 0));
        mma_commit(&empty[pp.st]);
        if (group_end) mma_commit(&dfull[buf]);
      }
      __syncwarp();
      if (group_end) { ++g; cg = 0; } else { ++cg; }
    }
  } else {
    // epilogue: lane quarter q of TMEM = rows q*32 .. q*32+31 of the tile; column half h
    const int q = warp & 3, h = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    float acc[GB_N / 2];
#pragma unroll
    for (int i = 0; i < GB_N / 2; ++i) acc[i] = 0.f;
    const int ngroups = (nk + GB_FLUSH - 1) / GB_FLUSH;
    for (int g = 0; g < ngroups; ++g) {
      const int buf = g & 1;
      mbar_wait(&dfull[buf], ((uint32_t)g >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < GB_N / 2; j += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + buf * GB_N + h * (GB_N / 2) + j, v);
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[j + e] += v[e];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dempty[buf]);
    }
    float* dst = p.part + ((int64_t)blockIdx.y * p.rows + (int64_t)blockIdx.x * TC_TILE + row) * GB_N + h * (GB_N / 2);
#pragma unroll
    for (int j = 0; j < GB_N / 2; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * GB_N);
}

// G[row, col0 + j] = sum_split part[split][row - row0][j]
__global__ void gram_bf16_reduce_kernel(const float* __restrict__ part, int splits, int64_t rows, int64_t row0, int col0,
                                        float* __restrict__ G, int64_t ldg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * (GB_N / 4)) return;
  const int64_t r = i / (GB_N / 4);
  const int j = (int)(i % (GB_N / 4)) * 4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < splits; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(part + ((int64_t)s * rows + r) * GB_N + j);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  *reinterpret_cast<float4*>(G + (row0 + r) * ldg + col0 + j) = a;
}

// out[t, s] = bf16 of the preprocessed value (zeros in the padding).  VEC: eight features per thread — two 16-byte
// loads, one 16-byte store (X 16-byte aligned with ldx % 4 == 0, out 16-byte aligned with ldo % 8 == 0); else one.
template <bool VEC>
__global__ void __launch_bounds__(256)
materialize_bf16_kernel(const float* __restrict__ X, int64_t T, int64_t S, int64_t ldx, const float* __restrict__ pivot,
                        const float* __restrict__ dscale, const float* __restrict__ ccorr,
                        const uint8_t* __restrict__ row_valid, int64_t rows_out, int64_t cols_out,
                        __nv_bfloat16* __restrict__ out, int64_t ldo) {
  constexpr int W = VEC ? 8 : 1;
  const int64_t s0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * W;
  if (s0 >= cols_out) return;
  float p[W], d[W], c[W];
#pragma unroll
  for (int e = 0; e < W; ++e) {
    const bool in = s0 + e < S;
    p[e] = in ? pivot[s0 + e] : 0.f;
    d[e] = in ? dscale[s0 + e] : 0.f;
    c[e] = (in && ccorr) ? ccorr[s0 + e] : 0.f;
  }
  const bool full = s0 + W <= S;  // (a vector never straddles cols_out: both are multiples of 8)
  for (int64_t t = blockIdx.y; t < rows_out; t += gridDim.y) {
    float v[W];
#pragma unroll
    for (int e = 0; e < W; ++e) v[e] = 0.f;
    if (t < T && (!row_valid || row_valid[t])) {
      float x[W];
      if (VEC && full) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(X + t * ldx + s0));
        const float4 b = __ldcs(reinterpret_cast<const float4*>(X + t * ldx + s0 + 4));
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
        if (W == 8) { x[W - 4] = b.x; x[W - 3] = b.y; x[W - 2] = b.z; x[W - 1] = b.w; }
      } else {
#pragma unroll
        for (int e = 0; e < W; ++e) x[e] = s0 + e < S ? X[t * ldx + s0 + e] : 0.f;
      }
#pragma unroll
      for (int e = 0; e < W; ++e) {
        const float y = x[e] - p[e];
        v[e] = s0 + e < S ? ((y == y) ? y * d[e] : 0.f) + c[e] : 0.f;
      }
    }
    if (VEC) {
      uint4 o;
      __nv_bfloat162 h;
      h = __floats2bfloat162_rn(v[0], v[1]); o.x = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2bfloat162_rn(v[2 % W], v[3 % W]); o.y = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2bfloat162_rn(v[4 % W], v[5 % W]); o.z = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2bfloat162_rn(v[6 % W], v[7 % W]); o.w = *reinterpret_cast<uint32_t*>(&h);
      *reinterpret_cast<uint4*>(out + t * ldo + s0) = o;
    } else {
      out[t * ldo + s0] = __float2bfloat16_rn(v[0]);
    }
  }
}

// project_tc.cu
typedef CUresult (*EncodeTiledFnB)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map_bf16(CUtensorMap* m, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer) {
  static EncodeTiledFnB enc = nullptr;
  if (!enc) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      enc = (EncodeTiledFnB)fp;
  }
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return XEOFS_E_UNSUPPORTED;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer}, gstr[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t bx[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer}, estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (bf16) failed (%d)", (int)r);
    return XEOFS_E_CUDA;
  }
  return XEOFS_OK;
}

static int gb_splits(int n_tiles, int ksteps) {
  // enough CTAs for two waves
  int64_t s = ceil_div(2 * (int64_t)num_sms(), n_tiles);
  if (s > ksteps) s = ksteps;
  return (int)(s < 1 ? 1 : s);
}

}  // namespace xb

using namespace xb;

extern "C" int xeofs_b200_materialize_bf16(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot,
                                           const float* dscale, const float* ccorr, const uint8_t* row_valid,
                                           int64_t rows_out, int64_t cols_out, void* out, int64_t ldo, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(X && pivot && dscale && out, "materialize_bf16: null pointer");
  XB_CHECK_ARG(T > 0 && S > 0 && ldx >= S && rows_out >= T && cols_out >= S && ldo >= cols_out, "materialize_bf16: bad shape");
  const bool vec = ldx % 4 == 0 && (uintptr_t)X % 16 == 0 && ldo % 8 == 0 && (uintptr_t)out % 16 == 0 && cols_out % 8 == 0;
  if (vec) {
    const unsigned gx = (unsigned)ceil_div(cols_out, 256 * 8);
    const unsigned gy = (unsigned)imin(rows_out, ceil_div(16 * (int64_t)num_sms(), gx) + 1);
    materialize_bf16_kernel<true><<<dim3(gx, gy), 256, 0, stream>>>(X, T, S, ldx, pivot, dscale, ccorr, row_valid, rows_out,
                                                                    cols_out, (__nv_bfloat16*)out, ldo);
  } else {
    materialize_bf16_kernel<false><<<dim3((unsigned)ceil_div(cols_out, 256), (unsigned)imin(rows_out, 4096)), 256, 0, stream>>>(
        X, T, S, ldx, pivot, dscale, ccorr, row_valid, rows_out, cols_out, (__nv_bfloat16*)out, ldo);
  }
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

extern "C" int64_t xeofs_b200_gram_rows_bf16_workspace_bytes(int64_t T_pad, int64_t S_pad) {
  const int ksteps = (int)(S_pad / GB_KS);
  int64_t mx = 0;
  for (int64_t c0 = 0; c0 < T_pad; c0 += GB_N) {
    const int n_tiles = (int)((T_pad - c0) / TC_TILE);
    const int64_t b = (int64_t)gb_splits(n_tiles, ksteps) * n_tiles * TC_TILE * GB_N * 4;
    if (b > mx) mx = b;
  }
  return mx + 256;
}

// G (T_pad x ldg fp32): the block-lower triangle (rows t >= 256 * (t' / 256)) of A A^T for the bf16 matrix A
// (T_pad x S_pad, T_pad a multiple of 256, S_pad a multiple of 64, row pitch ld elements).
extern "C" int xeofs_b200_gram_rows_bf16(const void* A, int64_t T_pad, int64_t S_pad, int64_t ld, float* G, int64_t ldg,
                                         void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(A && G && workspace, "gram_rows_bf16: null pointer");
  XB_CHECK_ARG(T_pad > 0 && T_pad % GB_N == 0 && S_pad > 0 && S_pad % GB_KS == 0 && ld >= S_pad && ld % 8 == 0 && ldg >= T_pad &&
                   ldg % 4 == 0 && ((uintptr_t)A % 128 == 0) && ((uintptr_t)G % 16 == 0) && ((uintptr_t)workspace % 16 == 0),
               "gram_rows_bf16: bad shape / alignment");
  XB_CHECK_ARG(workspace_bytes >= xeofs_b200_gram_rows_bf16_workspace_bytes(T_pad, S_pad), "gram_rows_bf16: workspace too small");
  CUtensorMap mA, mB;
  int rc = make_map_bf16(&mA, A, S_pad, T_pad, ld, GB_KS, TC_TILE);
  if (rc) return rc;
  rc = make_map_bf16(&mB, A, S_pad, T_pad, ld, GB_KS, GB_N);
  if (rc) return rc;
  const size_t smem = (size_t)GB_STAGES * (GB_A_BYTES + GB_B_BYTES) + 1024 + 256;
  XB_CUDA(cudaFuncSetAttribute(gram_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int ksteps = (int)(S_pad / GB_KS);
  for (int64_t c0 = 0; c0 < T_pad; c0 += GB_N) {
    const int n_tiles = (int)((T_pad - c0) / TC_TILE);
    const int splits = gb_splits(n_tiles, ksteps);
    GbParams p{};
    p.row_tile0 = (int)(c0 / TC_TILE);
    p.col0 = (int)c0;
    p.ksteps = ksteps;
    p.ksteps_per_split = (int)ceil_div(ksteps, splits);
    p.part = (float*)workspace;
    p.rows = (int64_t)n_tiles * TC_TILE;
    const int used = (int)ceil_div(ksteps, p.ksteps_per_split);
    gram_bf16_kernel<<<dim3((unsigned)n_tiles, (unsigned)used), 320, smem, stream>>>(mA, mB, p);
    XB_LAUNCH_CHECK();
    gram_bf16_reduce_kernel<<<(unsigned)ceil_div(p.rows * (GB_N / 4), 256), 256, 0, stream>>>((const float*)workspace, used, p.rows,
                                                                                             c0, (int)c0, G, ldg);
    XB_LAUNCH_CHECK();
  }
  return XEOFS_OK;
}
