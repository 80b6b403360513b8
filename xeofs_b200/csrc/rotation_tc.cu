// Varimax sweep on the tensor cores (R1; reference: linalg/_numpy/_rotation.py:162-177, one iteration of its loop).
//
// Per iteration the reference forms  B = X R  (S x m),  W = colsum(B^2)  and  X^H (B o (B^2 - W/S)).  Here the loadings
// Ln (Kaiser-normalised, space-side: row i = mode i, S contiguous) are streamed ONCE per iteration from HBM and both
// products run on tcgen05 (kind::tf32, hi/lo split on both operands = 3 MMAs per product, ~fp32 accuracy):
//
//   GEMM1   D1[j', s] = sum_i R[i, j'] Ln[i, s]         M = 128 (j'), N = 32 (s), K = i
//           A = R^T, hi and lo parts resident in TMEM (lane j', column i) for the whole kernel;
//           B = the Ln tile read MN-major (s contiguous).  kind::tf32 takes MN-major operands only in the
//           SWIZZLE_128B_BASE32B layout (32-byte chunks XOR-ed with row & 3, atoms of 4 K rows): the tile is fetched a
//           second time (an L2 hit) through a tensor map with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, which writes
//           exactly that image.  (With the plain 128-byte swizzle the tensor core returns zeros for this operand.)
//   stage   f = b^3 per TMEM lane (= per mode j'), W[j'] += b^2 — no cross-thread reduction — then f is split into
//           hi | lo and written straight back into TMEM as the A operand of GEMM2 (hi in place of D1).
//   GEMM2   G'[j', i] = sum_s f[j', s] Ln[i, s]          M = 128 (j'), N = nb (i), K = 32 (s)
//           A = f from TMEM, B = the tile as the 128-byte-swizzle tensor map delivered it, read K-major.
//
// B and f never leave the SM.  G' is a fresh TMEM accumulator per tile (the tensor core adds with truncation: short
// sums keep that noise random and far below the reference's stopping threshold of 1e-8 on sum(svals)); the epilogue
// warps add it into fp64 registers.
//
// One persistent CTA per SM, 10 warps: warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 lo-split of the tile (into
// both images), the b -> f stage and the fp64 accumulation.  MMA order G1(i), G2(i-1), G1(i+1), ... so that the stage
// of tile i runs under G2(i-1) and the accumulator flush of tile i-1 under G1(i+1).
#include "tc_common.cuh"

namespace xb {

constexpr int VT_TS = 32;         // features per tile
constexpr int VT_THREADS = 320;
constexpr int VT_MAX_STAGES = 4;
// TMEM columns
constexpr uint32_t VT_COL_D1 = 0;     // [2][32]  D1, then f_hi in place
constexpr uint32_t VT_COL_FLO = 64;   // [2][32]  f_lo
constexpr uint32_t VT_COL_G = 128;    // [<=128]  G' of the current tile
constexpr uint32_t VT_COL_RLO = 256;  // [k1]     lo part of R^T
constexpr uint32_t VT_COL_RHI = 384;  // [k1]     hi part of R^T

struct VtParams {
  int64_t S;
  int nb;         // N of GEMM2 = rows of a TMA box (m rounded up to 16)
  int ng;         // accumulator columns the epilogue sweeps (2 * NH >= nb)
  int k1;         // K of GEMM1 (m rounded up to 8)
  int stages;
  int ntiles;
  int x1;            // 1: single TF32 products (hi parts only, rounded to nearest) — the first phase of the iteration
  const float* rhi;  // [128][128]: rhi[j'][i] = TF32 bits of (float)R[i][j']
  const float* rlo;  // [128][128]: remainder
  double* gpart;     // [grid][128][ng]
  double* wpart;     // [grid][2][128]
};

// R (m x m fp64, row-major) -> A = R^T (row j', K = i):  hi = TF32 bits of (float)R, lo = remainder
__global__ void __launch_bounds__(256)
vt_prep_R_kernel(const double* __restrict__ R, int m, float* __restrict__ rhi, float* __restrict__ rlo, int x1) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 128 * 128; idx += gridDim.x * blockDim.x) {
    const int j = idx >> 7, k = idx & 127;
    const float r32 = (k < m && j < m) ? (float)R[(int64_t)k * m + j] : 0.f;
    // (single-product mode: the one part is the value rounded to nearest, not its upper bits)
    const float hi = x1 ? __uint_as_float(to_tf32(r32)) : __uint_as_float(__float_as_uint(r32) & 0xffffe000u);
    rhi[idx] = hi;
    // the tensor core reads the top 19 bits of a value: rounding the remainder to that width here (to nearest) keeps
    // the split unbiased, where the hardware's truncation would shrink every value by about 2^-22
    rlo[idx] = __uint_as_float(to_tf32(r32 - hi));
  }
}

// Gout[i, j'] (+)= sum_blocks gpart[b][j'][i];  Wout[j'] (+)= sum_blocks wpart[b][0..1][j']
__global__ void __launch_bounds__(256)
vt_reduce_kernel(const double* __restrict__ gpart, const double* __restrict__ wpart, int nblk, int m, int ng,
                 double* __restrict__ Gout, double* __restrict__ Wout, int accumulate) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < m * m) {
    const int j = idx / m, i = idx % m;
    double a = 0.0;
    for (int b = 0; b < nblk; ++b) a += gpart[((int64_t)b * 128 + j) * ng + i];
    double* o = &Gout[(int64_t)i * m + j];
    *o = accumulate ? *o + a : a;
  }
  if (idx < m) {
    double w = 0.0;
    for (int b = 0; b < nblk; ++b) w += wpart[(int64_t)b * 256 + idx] + wpart[(int64_t)b * 256 + 128 + idx];
    Wout[idx] = accumulate ? Wout[idx] + w : w;
  }
}

// MN-major operand tile of 32-bit values, SWIZZLE_128B_BASE32B: 32 values along M/N are contiguous (128 B), K rows
// 128 B apart, 32-byte chunks XOR-ed with (row & 3); atoms of 4 K rows `sbo_bytes` apart, groups of 32 along M/N
// `lbo_bytes` apart
__device__ __forceinline__ uint64_t make_mn32_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (1ull << 61);
}

template <int NH>  // accumulator columns per epilogue thread (2 * NH >= nb)
__global__ void __launch_bounds__(VT_THREADS, 1)
varimax_tc_kernel(const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapMN, const VtParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stages = p.stages, nb = p.nb, k1 = p.k1;
  const uint32_t boxb = (uint32_t)nb * 128;   // one TMA box: nb modes x 32 features
  const uint32_t stage_bytes = 4 * boxb;      // hi K-major image | hi MN-major image | lo K-major | lo MN-major
  uint64_t* bars = (uint64_t*)(smem + (size_t)stages * stage_bytes);
  uint64_t* full = bars;                        // [stages] TMA bytes landed
  uint64_t* loready = bars + VT_MAX_STAGES;     // [stages] lo parts of the tile written
  uint64_t* empty = bars + 2 * VT_MAX_STAGES;   // [stages] GEMM2 of the tile retired
  uint64_t* d1full = bars + 3 * VT_MAX_STAGES;  // [2] GEMM1 retired
  uint64_t* fready = d1full + 2;                // [2] f (hi | lo) in TMEM
  uint64_t* gfull = fready + 2;                 // G' holds the tile's product
  uint64_t* gdrained = gfull + 1;               // G' added into the registers
  uint64_t* rready = gdrained + 1;              // R^T (hi | lo) in TMEM
  uint32_t* tmem_slot = (uint32_t*)(rready + 1);

  const int nloc = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&loready[i], 8);
      mbar_init(&empty[i], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&d1full[b], 1);
      mbar_init(&fready[b], 8);
    }
    mbar_init(gfull, 1);
    mbar_init(gdrained, 8);
    mbar_init(rready, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    Pipe pp;
    for (int i = 0; i < nloc; ++i, pp.advance(stages)) {
      const int st = pp.st;
      mbar_wait(&empty[st], pp.ph ^ 1);
      if (elect_one()) {
        const int s0 = ((int)blockIdx.x + i * (int)gridDim.x) * VT_TS;
        uint8_t* dst = smem + (size_t)st * stage_bytes;
        mbar_expect_tx(&full[st], 2 * boxb);
        tma_load_2d(dst, &mapK, s0, 0, &full[st], HINT_EVICT_FIRST);
        tma_load_2d(dst + boxb, &mapMN, s0, 0, &full[st], HINT_EVICT_FIRST);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    const uint32_t idesc1 = make_idesc(VT_TS) | (1u << 16);  // B operand MN-major
    const uint32_t idesc2 = make_idesc(nb);
    const uint32_t smem_u = smem_u32(smem);
    constexpr uint32_t sbo1 = 512;  // atoms of 4 K rows x 128 B
    mbar_wait(rready, 0);
    tc_fence_after();
    Pipe p1, p2;  // stage / phase of the tile GEMM1 and GEMM2 work on
    for (int i = 0; i <= nloc; ++i) {
      if (i < nloc) {
        const int st = p1.st, b = i & 1;
        mbar_wait(&full[st], p1.ph);
        mbar_wait(&loready[st], p1.ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d1 = tmem_base + VT_COL_D1 + b * VT_TS;
          const uint32_t hi = smem_u + st * stage_bytes + boxb, lo = hi + 2 * boxb;  // the MN-major images
          // the small cross terms first: the tensor core adds with truncation, an error of up to one ulp of the
          // running sum per instruction, which costs nothing while the sum is 2^-11 of its final size
          if (!p.x1) {
            for (int kk = 0; kk < (k1 >> 3); ++kk) {
              const uint64_t b_hi = make_mn32_desc(hi + kk * 1024, boxb, sbo1);
              const uint64_t b_lo = make_mn32_desc(lo + kk * 1024, boxb, sbo1);
              mma_tf32_ts(d1, tmem_base + VT_COL_RLO + kk * 8, b_hi, idesc1, kk > 0);
              mma_tf32_ts(d1, tmem_base + VT_COL_RHI + kk * 8, b_lo, idesc1, 1);
            }
          }
          for (int kk = 0; kk < (k1 >> 3); ++kk)
            mma_tf32_ts(d1, tmem_base + VT_COL_RHI + kk * 8, make_mn32_desc(hi + kk * 1024, boxb, sbo1), idesc1,
                        !p.x1 || kk > 0);
          mma_commit(&d1full[b]);
        }
        __syncwarp();
        p1.advance(stages);
      }
      if (i >= 1) {
        const int j = i - 1, b = j & 1, st = p2.st;
        mbar_wait(&fready[b], ((uint32_t)j >> 1) & 1);
        if (j >= 1) mbar_wait(gdrained, (uint32_t)(j - 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t g = tmem_base + VT_COL_G;
          const uint32_t hi = smem_u + st * stage_bytes, lo = hi + 2 * boxb;  // the K-major images
          if (!p.x1) {
#pragma unroll
            for (int kk = 0; kk < VT_TS / 8; ++kk) {
              const uint32_t a_hi = tmem_base + VT_COL_D1 + b * VT_TS + kk * 8;
              const uint32_t a_lo = tmem_base + VT_COL_FLO + b * VT_TS + kk * 8;
              mma_tf32_ts(g, a_lo, make_b_desc(hi) + 2 * kk, idesc2, kk > 0);
              mma_tf32_ts(g, a_hi, make_b_desc(lo) + 2 * kk, idesc2, 1);
            }
          }
#pragma unroll
          for (int kk = 0; kk < VT_TS / 8; ++kk)
            mma_tf32_ts(g, tmem_base + VT_COL_D1 + b * VT_TS + kk * 8, make_b_desc(hi) + 2 * kk, idesc2, !p.x1 || kk > 0);
          mma_commit(&empty[st]);
          mma_commit(gfull);
        }
        __syncwarp();
        p2.advance(stages);
      }
    }
  } else {
    // ===================================================================== lo split, b -> f stage, fp64 accumulation
    const int q = warp & 3;         // TMEM lane quarter this warp may touch
    const int h = (warp - 2) >> 2;  // which half of the columns
    const int row = q * 32 + lane;  // mode j' = TMEM lane
    const int et = (warp - 2) * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t smem_u = smem_u32(smem);

    // R^T (hi | lo) -> TMEM (lane j', column i)
    for (int kc = h; kc * 8 < k1; kc += 2) {
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        const float* src = (part ? p.rlo : p.rhi) + row * 128 + kc * 8;
        const float4 a = *reinterpret_cast<const float4*>(src);
        const float4 c = *reinterpret_cast<const float4*>(src + 4);
        const uint32_t v[8] = {__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w),
                               __float_as_uint(c.x), __float_as_uint(c.y), __float_as_uint(c.z), __float_as_uint(c.w)};
        tmem_st8(tmem_base + lane_addr + (part ? VT_COL_RLO : VT_COL_RHI) + kc * 8, v);
      }
    }
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(rready);

    double acc[NH];
#pragma unroll
    for (int c = 0; c < NH; ++c) acc[c] = 0.0;
    double wacc = 0.0;

    // remainder of every value of the tile, into the K-major image (same place as in the hi image) and into the
    // MN-major one: the hi K-major image holds 16-byte chunk c of row r at c ^ (r & 7); the MN-major image holds
    // 32-byte chunk C at C ^ (r & 3), the two 16-byte halves in order
    auto split = [&](int st, uint32_t ph) {
      mbar_wait(&full[st], ph);
      const uint32_t hi = smem_u + st * stage_bytes, lo_k = hi + 2 * boxb, lo_mn = hi + 3 * boxb;
      for (uint32_t off = et * 16; off < boxb && !p.x1; off += 256 * 16) {
        const float4 v = lds128(hi + off);
        float4 r;
        r.x = __uint_as_float(to_tf32(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u)));
        r.y = __uint_as_float(to_tf32(v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u)));
        r.z = __uint_as_float(to_tf32(v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u)));
        r.w = __uint_as_float(to_tf32(v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u)));
        sts128(lo_k + off, r);
        const uint32_t rr = off >> 7, c = ((off >> 4) & 7) ^ (rr & 7);
        const uint32_t pos = (((c >> 1) ^ (rr & 3)) << 1) | (c & 1);
        sts128(lo_mn + (rr << 7) + (pos << 4), r);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&loready[st]);
    };
    auto stage_f = [&](int i) {
      const int b = i & 1;
      mbar_wait(&d1full[b], ((uint32_t)i >> 1) & 1);
      tc_fence_after();
      const uint32_t col = b * VT_TS + h * 16;
      float v[16];
      tmem_ld16(tmem_base + lane_addr + VT_COL_D1 + col, v);
      float w2 = 0.f;
      uint32_t fh[16], fl[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float b2 = v[e] * v[e];
        w2 += b2;
        const float f = b2 * v[e];
        fh[e] = p.x1 ? to_tf32(f) : (__float_as_uint(f) & 0xffffe000u);
        fl[e] = to_tf32(f - __uint_as_float(fh[e]));
      }
      tmem_st16(tmem_base + lane_addr + VT_COL_D1 + col, fh);
      if (!p.x1) tmem_st16(tmem_base + lane_addr + VT_COL_FLO + col, fl);
      wacc += (double)w2;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&fready[b]);
    };
    auto flush = [&](int j) {
      mbar_wait(gfull, (uint32_t)j & 1);
      tc_fence_after();
#pragma unroll
      for (int gi = 0; gi < NH / 8; ++gi) {
        float v[8];
        tmem_ld8(tmem_base + lane_addr + VT_COL_G + h * NH + gi * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[gi * 8 + e] += (double)v[e];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(gdrained);
    };

    Pipe ps;  // stage / phase of the next tile to split
    split(ps.st, ps.ph);
    ps.advance(stages);
    for (int i = 0; i < nloc; ++i) {
      if (i + 1 < nloc) {
        split(ps.st, ps.ph);
        ps.advance(stages);
      }
      stage_f(i);
      if (i >= 1) flush(i - 1);
    }
    flush(nloc - 1);

    double* gp = p.gpart + ((int64_t)blockIdx.x * 128 + row) * (2 * NH) + h * NH;
#pragma unroll
    for (int c = 0; c < NH; c += 2) *reinterpret_cast<double2*>(gp + c) = make_double2(acc[c], acc[c + 1]);
    p.wpart[(int64_t)blockIdx.x * 256 + h * 128 + row] = wacc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static int vt_n2(int64_t m) {  // instantiated accumulator widths
  return m <= 32 ? 32 : m <= 64 ? 64 : m <= 96 ? 96 : m <= 112 ? 112 : 128;
}

int64_t varimax_tc_workspace_bytes(int64_t S, int64_t m) {
  (void)S;
  const int64_t ng = vt_n2(m);
  return 65536 /*R^T hi*/ + 65536 /*R^T lo*/ + (int64_t)num_sms() * (128 * ng + 256) * 8 + 1024;
}

bool varimax_tc_supported(const float* L, int64_t S, int64_t m, int64_t ld) {
  return m >= 2 && m <= 128 && S >= 1 && ld % 4 == 0 && ((uintptr_t)L % 16 == 0) && tensor_maps_available() &&
         S + 64 < (int64_t)1 << 31;
}

int varimax_sweep_tc(const float* L, int64_t S, int64_t m, int64_t ld, const double* R, double* Gout, double* Wout,
                     int accumulate, int products, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (workspace_bytes < varimax_tc_workspace_bytes(S, m)) {
    set_error("varimax_sweep: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes,
              (long long)varimax_tc_workspace_bytes(S, m));
    return XEOFS_E_WORKSPACE;
  }
  const int ng = vt_n2(m), k1 = (int)round_up(m, 8);
  const int nb = (int)lpad(m);  // rows the caller's space-side matrix holds (pad rows zero)
  VtParams p{};
  p.S = S; p.nb = nb; p.ng = ng; p.k1 = k1;
  p.x1 = products == 1 ? 1 : 0;
  p.ntiles = (int)ceil_div(S, VT_TS);
  const int stage_bytes = nb * 512;
  const int budget = 227 * 1024 - 1024 /*alignment*/ - 512 /*barriers*/;
  int stages = budget / stage_bytes;
  if (stages > VT_MAX_STAGES) stages = VT_MAX_STAGES;
  const int forced = env_int("XEOFS_VT_STAGES", 0);
  if (forced >= 2 && forced < stages) stages = forced;
  if (stages < 2) {
    set_error("varimax_sweep: no pipeline shape fits m=%lld", (long long)m);
    return XEOFS_E_UNSUPPORTED;
  }
  p.stages = stages;
  uint8_t* ws = (uint8_t*)workspace;
  float* rhi = (float*)ws; ws += 65536;
  float* rlo = (float*)ws; ws += 65536;
  const int grid = (int)imin(num_sms(), p.ntiles);
  double* gpart = (double*)ws; ws += (int64_t)grid * 128 * ng * 8;
  double* wpart = (double*)ws;
  p.rhi = rhi; p.rlo = rlo; p.gpart = gpart; p.wpart = wpart;
  vt_prep_R_kernel<<<16, 256, 0, stream>>>(R, (int)m, rhi, rlo, p.x1);
  XB_LAUNCH_CHECK();
  CUtensorMap mapK, mapMN;
  int rc = make_map2(&mapK, L, S, nb, ld, VT_TS, nb, 1);
  if (rc) return rc;
  rc = make_map2(&mapMN, L, S, nb, ld, VT_TS, nb, 2);
  if (rc) return rc;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 512;
#define XB_VT(NHV)                                                                                                     \
  XB_CUDA(cudaFuncSetAttribute(varimax_tc_kernel<NHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
  varimax_tc_kernel<NHV><<<grid, VT_THREADS, smem, stream>>>(mapK, mapMN, p)
  switch (ng) {
    case 32: XB_VT(16); break;
    case 64: XB_VT(32); break;
    case 96: XB_VT(48); break;
    case 112: XB_VT(56); break;
    default: XB_VT(64); break;
  }
#undef XB_VT
  XB_LAUNCH_CHECK();
  vt_reduce_kernel<<<(unsigned)ceil_div(m * m, 256), 256, 0, stream>>>(gpart, wpart, grid, (int)m, ng, Gout, Wout, accumulate);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

}  // namespace xb
