// Varimax sweep on the tensor cores (R1; reference: linalg/_numpy/_rotation.py:162-177, one iteration of its loop).
//
// Per iteration the reference forms  B = X R  (S x m),  W = colsum(B^2)  and  X^H (B o (B^2 - W/S)).  Here the loadings
// Ln (Kaiser-normalised, space-side: row i = mode i, S contiguous) are streamed ONCE per iteration and both products
// run on tcgen05 (kind::tf32, hi/lo split on both operands = 3 MMAs per product, ~fp32 accuracy):
//
//   GEMM1   D1[j', s] = sum_i R[i, j'] Ln[i, s]         M = 128 (j'), N = 64 (s), K = i
//           A = R^T: hi part from shared memory (K-major image, loaded once per CTA), lo part from TMEM;
//           B = the Ln tile as TMA delivered it (box = [n2 modes][32 s], 128-byte swizzle), read MN-major.
//   stage   f = b^3 per TMEM lane (= per mode j'), W[j'] += b^2 — no cross-thread reduction — then f is split into
//           hi | lo and written straight back into TMEM as the A operand of GEMM2 (hi in place of D1).
//   GEMM2   G'[j', i] = sum_s f[j', s] Ln[i, s]          M = 128 (j'), N = n2 (i), K = 64 (s)
//           A = f from TMEM, B = the SAME shared-memory tile, now read K-major.  (n2 = m rounded up to 16)
//
// So the tile goes HBM -> shared memory once and is consumed by both products; B and f never leave the SM.  G' is a
// fresh TMEM accumulator per tile (the tensor core adds with truncation: short sums keep that noise random and far
// below the reference's stopping threshold of 1e-8 on sum(svals)); the epilogue warps add it into fp64 registers.
//
// One persistent CTA per SM, 10 warps: warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 lo-split of the tile,
// the b -> f stage and the fp64 accumulation.  MMA order G1(i), G2(i-1), G1(i+1), ... so that the stage of tile i runs
// under G2(i-1) and the accumulator flush of tile i-1 under G1(i+1).
#include "tc_common.cuh"

namespace xb {

constexpr int VT_TS = 64;         // features per tile
constexpr int VT_THREADS = 320;
constexpr int VT_MAX_STAGES = 4;
// TMEM columns
constexpr uint32_t VT_COL_D1 = 0;     // [2][64]  D1, then f_hi in place
constexpr uint32_t VT_COL_FLO = 128;  // [2][64]  f_lo
constexpr uint32_t VT_COL_G = 256;    // [<=128]  G' of the current tile
constexpr uint32_t VT_COL_RLO = 384;  // [k1]     lo part of R^T

struct VtParams {
  int64_t S;
  int nb;         // N of GEMM2 = rows of a TMA box (m rounded up to 16)
  int ng;         // accumulator columns the epilogue sweeps (2 * NH >= nb)
  int k1;         // K of GEMM1 (m rounded up to 8)
  int stages;
  int ntiles;
  int rhi_sw128;  // layout of the R^T hi image: 1 = 32-wide K slabs, SWIZZLE_128B; 0 = 8-wide K blocks without swizzle
  int rhi_bytes;
  int flags;      // debug: 1 swaps LBO/SBO of the un-swizzled descriptor, 2 of the MN-major descriptor
  const float* rhi_img;
  const float* rlo;  // [128][128]: rlo[j'][i]
  double* gpart;     // [grid][128][ng]
  double* wpart;     // [grid][2][128]
};

// R (m x m fp64, row-major) -> operand images of A = R^T (row j', K = i):  hi = TF32 bits of (float)R, lo = remainder
__global__ void __launch_bounds__(256)
vt_prep_R_kernel(const double* __restrict__ R, int m, int k1, int sw128, float* __restrict__ rhi_img, float* __restrict__ rlo) {
  const int kmax = sw128 ? (k1 + 31) / 32 * 32 : k1;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 128 * 128; idx += gridDim.x * blockDim.x) {
    const int j = idx >> 7, k = idx & 127;
    const float r32 = (k < m && j < m) ? (float)R[(int64_t)k * m + j] : 0.f;
    const float hi = __uint_as_float(__float_as_uint(r32) & 0xffffe000u);
    rlo[idx] = r32 - hi;
    if (k < kmax) {
      const int off = sw128 ? (k >> 5) * 4096 + img_offset(j, k & 31)
                            : (k >> 3) * 1024 + (j >> 3) * 64 + ((k & 7) >> 2) * 32 + (j & 7) * 4 + (k & 3);
      rhi_img[off] = hi;
    }
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

template <int NH>  // accumulator columns per epilogue thread (2 * NH >= nb)
__global__ void __launch_bounds__(VT_THREADS, 1)
varimax_tc_kernel(const __grid_constant__ CUtensorMap mapL, const VtParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stages = p.stages, nb = p.nb, k1 = p.k1;
  const uint32_t boxb = (uint32_t)nb * 128;   // one TMA box: nb modes x 32 features
  const uint32_t stage_bytes = 4 * boxb;      // hi box 0 | hi box 1 | lo box 0 | lo box 1
  uint8_t* rhi = smem + (size_t)stages * stage_bytes;
  uint64_t* bars = (uint64_t*)(rhi + p.rhi_bytes);
  uint64_t* full = bars;                        // [stages] TMA bytes landed
  uint64_t* loready = bars + VT_MAX_STAGES;     // [stages] lo part of the tile written
  uint64_t* empty = bars + 2 * VT_MAX_STAGES;   // [stages] GEMM2 of the tile retired
  uint64_t* d1full = bars + 3 * VT_MAX_STAGES;  // [2] GEMM1 retired
  uint64_t* fready = d1full + 2;                // [2] f (hi | lo) in TMEM
  uint64_t* gfull = fready + 2;                 // G' holds the tile's product
  uint64_t* gdrained = gfull + 1;               // G' added into the registers
  uint64_t* rfull = gdrained + 1;               // R^T hi image in shared memory
  uint64_t* rloready = rfull + 1;               // R^T lo part in TMEM
  uint32_t* tmem_slot = (uint32_t*)(rloready + 1);

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
    mbar_init(rfull, 1);
    mbar_init(rloready, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      mbar_expect_tx(rfull, (uint32_t)p.rhi_bytes);
      for (int o = 0; o < p.rhi_bytes; o += 4096) bulk_load_1d(rhi + o, (const uint8_t*)p.rhi_img + o, 4096, rfull);
    }
    __syncwarp();
    Pipe pp;
    for (int i = 0; i < nloc; ++i, pp.advance(stages)) {
      const int st = pp.st;
      mbar_wait(&empty[st], pp.ph ^ 1);
      if (elect_one()) {
        const int s0 = ((int)blockIdx.x + i * (int)gridDim.x) * VT_TS;
        uint8_t* dst = smem + (size_t)st * stage_bytes;
        mbar_expect_tx(&full[st], 2 * boxb);
        tma_load_2d(dst, &mapL, s0, 0, &full[st], HINT_EVICT_FIRST);
        tma_load_2d(dst + boxb, &mapL, s0 + 32, 0, &full[st], HINT_EVICT_FIRST);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    const uint32_t idesc1 = make_idesc(VT_TS) | (1u << 16);  // B operand MN-major
    const uint32_t idesc2 = make_idesc(nb);
    const uint32_t rhi_u32 = smem_u32(rhi), smem_u = smem_u32(smem);
    const bool swap_ns = p.flags & 1, swap_mn = p.flags & 2;
    mbar_wait(rfull, 0);
    mbar_wait(rloready, 0);
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
          const uint32_t hi = smem_u + st * stage_bytes, lo = hi + 2 * boxb;
          for (int kk = 0; kk < (k1 >> 3); ++kk) {
            const uint64_t a_hi = p.rhi_sw128 ? make_b_desc(rhi_u32 + (kk >> 2) * 16384) + 2 * (kk & 3)
                                  : swap_ns   ? make_nosw_desc(rhi_u32 + kk * 4096, 256, 128)
                                              : make_nosw_desc(rhi_u32 + kk * 4096, 128, 256);
            const uint64_t b_hi = swap_mn ? make_mn_desc(hi + kk * 1024, 1024, boxb) : make_mn_desc(hi + kk * 1024, boxb, 1024);
            const uint64_t b_lo = swap_mn ? make_mn_desc(lo + kk * 1024, 1024, boxb) : make_mn_desc(lo + kk * 1024, boxb, 1024);
            mma_tf32_ss(d1, a_hi, b_hi, idesc1, kk > 0);
            mma_tf32_ts(d1, tmem_base + VT_COL_RLO + kk * 8, b_hi, idesc1, 1);
            mma_tf32_ss(d1, a_hi, b_lo, idesc1, 1);
          }
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
          const uint32_t hi = smem_u + st * stage_bytes, lo = hi + 2 * boxb;
#pragma unroll
          for (int kk = 0; kk < VT_TS / 8; ++kk) {
            const uint32_t a_hi = tmem_base + VT_COL_D1 + b * VT_TS + kk * 8;
            const uint32_t a_lo = tmem_base + VT_COL_FLO + b * VT_TS + kk * 8;
            const uint64_t b_hi = make_b_desc(hi + (kk >> 2) * boxb) + 2 * (kk & 3);
            const uint64_t b_lo = make_b_desc(lo + (kk >> 2) * boxb) + 2 * (kk & 3);
            mma_tf32_ts(g, a_hi, b_hi, idesc2, kk > 0);
            mma_tf32_ts(g, a_lo, b_hi, idesc2, 1);
            mma_tf32_ts(g, a_hi, b_lo, idesc2, 1);
          }
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

    // lo part of R^T -> TMEM (lane j', column i)
    for (int kc = h; kc * 8 < k1; kc += 2) {
      const float4 a = *reinterpret_cast<const float4*>(p.rlo + row * 128 + kc * 8);
      const float4 c = *reinterpret_cast<const float4*>(p.rlo + row * 128 + kc * 8 + 4);
      const uint32_t v[8] = {__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w),
                             __float_as_uint(c.x), __float_as_uint(c.y), __float_as_uint(c.z), __float_as_uint(c.w)};
      tmem_st8(tmem_base + lane_addr + VT_COL_RLO + kc * 8, v);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(rloready);

    double acc[NH];
#pragma unroll
    for (int c = 0; c < NH; ++c) acc[c] = 0.0;
    double wacc = 0.0;

    auto split = [&](int st, uint32_t ph) {
      mbar_wait(&full[st], ph);
      const uint32_t hi = smem_u + st * stage_bytes, lo = hi + 2 * boxb;
      for (uint32_t off = et * 16; off < 2 * boxb; off += 256 * 16) {
        const float4 v = lds128(hi + off);
        float4 r;
        r.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
        r.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        r.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
        r.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
        sts128(lo + off, r);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&loready[st]);
    };
    auto stage_f = [&](int i) {
      const int b = i & 1;
      mbar_wait(&d1full[b], ((uint32_t)i >> 1) & 1);
      tc_fence_after();
      float w2 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t col = b * VT_TS + h * 32 + c * 16;
        float v[16];
        tmem_ld16(tmem_base + lane_addr + VT_COL_D1 + col, v);
        uint32_t fh[16], fl[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float b2 = v[e] * v[e];
          w2 += b2;
          const float f = b2 * v[e];
          fh[e] = __float_as_uint(f) & 0xffffe000u;
          fl[e] = __float_as_uint(f - __uint_as_float(fh[e]));
        }
        tmem_st16(tmem_base + lane_addr + VT_COL_D1 + col, fh);
        tmem_st16(tmem_base + lane_addr + VT_COL_FLO + col, fl);
      }
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
  return 65536 /*R^T hi image*/ + 65536 /*R^T lo*/ + (int64_t)num_sms() * (128 * ng + 256) * 8 + 1024;
}

bool varimax_tc_supported(const float* L, int64_t S, int64_t m, int64_t ld) {
  return m >= 2 && m <= 128 && S >= 1 && ld % 4 == 0 && ((uintptr_t)L % 16 == 0) && tensor_maps_available() &&
         S + 64 < (int64_t)1 << 31;
}

int varimax_sweep_tc(const float* L, int64_t S, int64_t m, int64_t ld, const double* R, double* Gout, double* Wout,
                     int accumulate, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (workspace_bytes < varimax_tc_workspace_bytes(S, m)) {
    set_error("varimax_sweep: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes,
              (long long)varimax_tc_workspace_bytes(S, m));
    return XEOFS_E_WORKSPACE;
  }
  const int ng = vt_n2(m), k1 = (int)round_up(m, 8);
  const int nb = (int)lpad(m);  // rows the caller's space-side matrix holds (pad rows zero)
  VtParams p{};
  p.S = S; p.nb = nb; p.ng = ng; p.k1 = k1;
  p.ntiles = (int)ceil_div(S, VT_TS);
  p.rhi_sw128 = env_int("XEOFS_VT_RHI_SW128", 0) ? 1 : 0;
  p.flags = env_int("XEOFS_VT_FLAGS", 0);
  p.rhi_bytes = p.rhi_sw128 ? (int)ceil_div(k1, 32) * 16384 : (k1 / 8) * 4096;
  const int stage_bytes = nb * 512;
  const int budget = 227 * 1024 - 1024 /*alignment*/ - 512 /*barriers*/ - p.rhi_bytes;
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
  float* rhi_img = (float*)ws; ws += 65536;
  float* rlo = (float*)ws; ws += 65536;
  const int grid = (int)imin(num_sms(), p.ntiles);
  double* gpart = (double*)ws; ws += (int64_t)grid * 128 * ng * 8;
  double* wpart = (double*)ws;
  p.rhi_img = rhi_img; p.rlo = rlo; p.gpart = gpart; p.wpart = wpart;
  XB_CUDA(cudaMemsetAsync(rhi_img, 0, 65536, stream));
  vt_prep_R_kernel<<<16, 256, 0, stream>>>(R, (int)m, k1, p.rhi_sw128, rhi_img, rlo);
  XB_LAUNCH_CHECK();
  CUtensorMap mapL;
  int rc = make_map2(&mapL, L, S, nb, ld, 32, nb, true);
  if (rc) return rc;
  const size_t smem = (size_t)stages * stage_bytes + p.rhi_bytes + 1024 + 512;
#define XB_VT(NHV)                                                                                                     \
  XB_CUDA(cudaFuncSetAttribute(varimax_tc_kernel<NHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
  varimax_tc_kernel<NHV><<<grid, VT_THREADS, smem, stream>>>(mapL, p)
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
