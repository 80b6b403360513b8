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

// Gout[i, j'] (+)= sum_blocks gpart[b][j'][i];  Wout[j'] (+)= sum_blocks wpart[b][0..wparts-1][j'].  One warp per four
// consecutive entries: the lanes share the blocks (8 lanes per entry, every lane a fixed subset in a fixed order, then a
// shuffle tree — the result does not depend on scheduling).
__global__ void __launch_bounds__(256)
vt_reduce_kernel(const double* __restrict__ gpart, const double* __restrict__ wpart, int nblk, int m, int ng,
                 double* __restrict__ Gout, double* __restrict__ Wout, int accumulate, int wparts) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int sub = lane >> 2, e = lane & 3;  // 8 subsets of the blocks x 4 entries
  const int idx = warp * 4 + e;
  const int n_g = m * m;
  if (warp * 4 >= n_g + m) return;
  double a = 0.0;
  if (idx < n_g) {
    const int j = idx / m, i = idx % m;
    for (int b = sub; b < nblk; b += 8) a += gpart[((int64_t)b * 128 + j) * ng + i];
  } else if (idx < n_g + m) {
    const int j = idx - n_g;
    for (int b = sub; b < nblk; b += 8)
      for (int h = 0; h < wparts; ++h) a += wpart[((int64_t)b * wparts + h) * 128 + j];
  }
  a += __shfl_xor_sync(0xffffffffu, a, 4);
  a += __shfl_xor_sync(0xffffffffu, a, 8);
  a += __shfl_xor_sync(0xffffffffu, a, 16);
  if (sub == 0) {
    if (idx < n_g) {
      double* o = &Gout[(int64_t)(idx % m) * m + idx / m];
      *o = accumulate ? *o + a : a;
    } else if (idx < n_g + m) {
      Wout[idx - n_g] = accumulate ? Wout[idx - n_g] + a : a;
    }
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

// ------------------------------------------------------------------------------------------------
// The same sweep on tiles of 64 features, fed from a PACKED copy of the loadings.
//
// Why 64: a kind::tf32 MMA with M = 128, K = 8 costs ~64 cycles of tensor-pipe issue whether N is 32 or 64
// (profiles/r01_mma_issue_probe.txt), and GEMM1 is 3 * k1 / 8 of them per tile: at N = 32 it alone is 2500 cycles per
// 32 features for 100 modes, the whole sweep 3400.  With N = 64 the same instructions cover twice the features.
// Why packed: the loadings do not change over the ~100 iterations of a rotation.  Fetching a tile as 4 x nb rows of
// 128 bytes through two tensor maps caps the sweep near 1 ms (1.6 TB/s) whatever the arithmetic; one pass
// (vt_pack_kernel) therefore rewrites them tile by tile as the K-major SWIZZLE_128B image the tensor core reads — two
// boxes of nb rows x 32 features, 2 nb 128 contiguous bytes — and every tile arrives as ONE bulk copy.
//   * the epilogue warps derive the MN-major image (same rows, 32-byte chunks XOR-ed with row & 3) and, in the
//     three-product mode, both remainder images in shared memory;
//   * GEMM1 reads the two boxes as one MN-major operand (the second group of 32 features LBO = one box further),
//     GEMM2 runs K = 64 (four K steps per box);
//   * the K-major images live from their arrival to the end of GEMM2 (three slots or four), the MN-major ones only
//     during GEMM1 (one slot, rewritten while GEMM2 of the previous tile runs; two where shared memory has room — the
//     single-product mode — so that GEMM1 of the next tile never waits for the epilogue warps);
//   * 16 epilogue warps (lane quarter x column quarter); f_lo is single-buffered in TMEM (512 columns: D1 2 x 64,
//     f_lo 64, G' ng, R^T hi and lo k1 each) and written once GEMM2 of the previous tile has retired.
constexpr int V2_TS = 64;
constexpr int V2_THREADS = 576;
constexpr int V2_MAX_SLOTS = 4;

struct Vt2Params {
  int nb, ng, k1;
  int kslots;   // K-major slots (a tile's image from its arrival to the end of GEMM2)
  int mnslots;  // MN-major slots (1 or 2)
  int ntiles;   // of V2_TS features
  int x1;
  int pf;       // tiles ahead whose bytes the producer asks into L2
  uint32_t col_flo, col_g, col_rlo, col_rhi;  // TMEM columns (D1 = [2][64] at 0)
  const float* packed;  // [ntiles][2][nb][32]
  const float* rhi;
  const float* rlo;
  double* gpart;  // [grid][128][ng]
  double* wpart;  // [grid][4][128]
};

// Ln (space-side, nb rows of S features, pad rows zero) -> packed[tile][box][row][32]: 16-byte chunk c of a row
// stored at c ^ (row & 7), zeros beyond S.  One thread per chunk.
__global__ void __launch_bounds__(256)
vt_pack_kernel(const float* __restrict__ L, int64_t S, int64_t ld, int nb, int64_t ntiles, float* __restrict__ packed) {
  const int64_t total = ntiles * nb * 16;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(idx & 15);  // chunk inside the 256 bytes of (tile, row): box = cc >> 3, c = cc & 7
    const int64_t tr = idx >> 4;
    const int row = (int)(tr % nb);
    const int64_t tile = tr / nb;
    const int64_t s = tile * V2_TS + cc * 4;
    const float* src = L + (int64_t)row * ld + s;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s + 3 < S) v = *reinterpret_cast<const float4*>(src);
    else {
      if (s < S) v.x = src[0];
      if (s + 1 < S) v.y = src[1];
      if (s + 2 < S) v.z = src[2];
    }
    const int box = cc >> 3, c = cc & 7;
    float* dst = packed + ((tile * 2 + box) * nb + row) * 32 + ((c ^ (row & 7)) << 2);
    *reinterpret_cast<float4*>(dst) = v;
  }
}

template <int NH>  // accumulator columns per epilogue thread (4 * NH = ng)
__global__ void __launch_bounds__(V2_THREADS, 1)
varimax_tc2_kernel(const Vt2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kslots = p.kslots, nb = p.nb, k1 = p.k1;
  const bool x1 = p.x1 != 0;
  const uint32_t boxb = (uint32_t)nb * 128;        // one box: nb modes x 32 features
  const uint32_t imgb = 2 * boxb;                  // one image of a tile: two boxes
  const uint32_t slotb = (x1 ? 1u : 2u) * imgb;    // hi image | lo image
  const uint32_t mn_off = (uint32_t)kslots * slotb;  // the MN-major slots follow the K-major ones
  const bool mn2 = p.mnslots == 2;
  uint64_t* bars = (uint64_t*)(smem + (size_t)(kslots + p.mnslots) * slotb);
  uint64_t* fullK = bars;                        // [kslots] the tile's bytes landed
  uint64_t* emptyK = bars + V2_MAX_SLOTS;        // [kslots] GEMM2 of the tile retired
  uint64_t* d1full = bars + 2 * V2_MAX_SLOTS;    // [2] GEMM1 retired (D1 full, its MN-major slot free)
  uint64_t* fready = d1full + 2;                 // [2] f (hi | lo) in TMEM
  uint64_t* mnready = fready + 2;                // [2] the derived images of a tile are written
  uint64_t* gfull = mnready + 2;                 // G' holds the tile's product
  uint64_t* gdrained = gfull + 1;                // G' read into registers
  uint64_t* rready = gdrained + 1;               // R^T (hi | lo) in TMEM
  uint32_t* tmem_slot = (uint32_t*)(rready + 1);

  const int nloc = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kslots; ++i) {
      mbar_init(&fullK[i], 1);
      mbar_init(&emptyK[i], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&d1full[b], 1);
      mbar_init(&fready[b], 16);
      mbar_init(&mnready[b], 16);
    }
    mbar_init(gfull, 1);
    mbar_init(gdrained, 16);
    mbar_init(rready, 16);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== producer: one bulk copy per tile
    Pipe pp;
    for (int i = 0; i < nloc; ++i, pp.advance(kslots)) {
      const int st = pp.st;
      mbar_wait(&emptyK[st], pp.ph ^ 1);
      if (elect_one()) {
        const int64_t tile = (int64_t)blockIdx.x + (int64_t)i * gridDim.x;
        mbar_expect_tx(&fullK[st], imgb);
        bulk_load_1d_hint(smem + (size_t)st * slotb, p.packed + tile * (imgb / 4), imgb, &fullK[st], HINT_EVICT_FIRST);
        if (p.pf > 0 && i + p.pf < nloc)  // the slot ring is short: have the tiles after the next ones wait in L2
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.packed + (tile + (int64_t)p.pf * gridDim.x) * (imgb / 4)),
                       "r"(imgb)
                       : "memory");
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    const uint32_t idesc1 = make_idesc(V2_TS) | (1u << 16);  // B operand MN-major
    const uint32_t idesc2 = make_idesc(nb);
    const uint32_t smem_u = smem_u32(smem);
    constexpr uint32_t sbo1 = 512;  // atoms of 4 K rows x 128 B
    const uint32_t RHI = tmem_base + p.col_rhi, RLO = tmem_base + p.col_rlo;
    mbar_wait(rready, 0);
    tc_fence_after();
    Pipe p2;  // slot / phase of the tile GEMM2 works on
    for (int i = 0; i <= nloc; ++i) {
      if (i < nloc) {
        const int b = i & 1;
        // (one MN-major slot: barrier 0 completes once per tile; two: barrier b once per two tiles)
        if (mn2) mbar_wait(&mnready[b], ((uint32_t)i >> 1) & 1);
        else mbar_wait(&mnready[0], (uint32_t)i & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d1 = tmem_base + b * V2_TS;
          const uint32_t hi = smem_u + mn_off + (mn2 ? (uint32_t)b * slotb : 0u), lo = hi + imgb;
          if (!x1) {  // the small cross terms first (see above)
            for (int kk = 0; kk < (k1 >> 3); ++kk) {
              mma_tf32_ts(d1, RLO + kk * 8, make_mn32_desc(hi + kk * 1024, boxb, sbo1), idesc1, kk > 0);
              mma_tf32_ts(d1, RHI + kk * 8, make_mn32_desc(lo + kk * 1024, boxb, sbo1), idesc1, 1);
            }
          }
          for (int kk = 0; kk < (k1 >> 3); ++kk)
            mma_tf32_ts(d1, RHI + kk * 8, make_mn32_desc(hi + kk * 1024, boxb, sbo1), idesc1, !x1 || kk > 0);
          mma_commit(&d1full[b]);
        }
        __syncwarp();
      }
      if (i >= 1) {
        const int j = i - 1, b = j & 1, st = p2.st;
        mbar_wait(&fready[b], ((uint32_t)j >> 1) & 1);
        if (j >= 1) mbar_wait(gdrained, (uint32_t)(j - 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t g = tmem_base + p.col_g;
          const uint32_t hi = smem_u + st * slotb, lo = hi + imgb;  // (their arrival is implied by mnready of tile j)
          const uint32_t a_hi = tmem_base + b * V2_TS, a_lo = tmem_base + p.col_flo;
          if (!x1) {
#pragma unroll
            for (int kk = 0; kk < V2_TS / 8; ++kk) {
              const uint32_t off = (kk >> 2) * boxb;
              mma_tf32_ts(g, a_lo + kk * 8, make_b_desc(hi + off) + 2 * (kk & 3), idesc2, kk > 0);
              mma_tf32_ts(g, a_hi + kk * 8, make_b_desc(lo + off) + 2 * (kk & 3), idesc2, 1);
            }
          }
#pragma unroll
          for (int kk = 0; kk < V2_TS / 8; ++kk)
            mma_tf32_ts(g, a_hi + kk * 8, make_b_desc(hi + (kk >> 2) * boxb) + 2 * (kk & 3), idesc2, !x1 || kk > 0);
          mma_commit(&emptyK[st]);
          mma_commit(gfull);
        }
        __syncwarp();
        p2.advance(kslots);
      }
    }
  } else {
    // ===================================================================== derived images, b -> f stage, fp64 accumulation
    const int q = warp & 3;         // TMEM lane quarter this warp may touch
    const int c = (warp - 2) >> 2;  // which quarter of the columns
    const int row = q * 32 + lane;  // mode j' = TMEM lane
    const int et = (warp - 2) * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t smem_u = smem_u32(smem);

    // R^T (hi | lo) -> TMEM (lane j', column i)
    for (int kc = c; kc * 8 < k1; kc += 4) {
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        if (part && x1) continue;
        const float* src = (part ? p.rlo : p.rhi) + row * 128 + kc * 8;
        const float4 a = *reinterpret_cast<const float4*>(src);
        const float4 d = *reinterpret_cast<const float4*>(src + 4);
        const uint32_t v[8] = {__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w),
                               __float_as_uint(d.x), __float_as_uint(d.y), __float_as_uint(d.z), __float_as_uint(d.w)};
        tmem_st8(tmem_base + lane_addr + (part ? p.col_rlo : p.col_rhi) + kc * 8, v);
      }
    }
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(rready);

    double acc[NH];
#pragma unroll
    for (int e = 0; e < NH; ++e) acc[e] = 0.0;
    double wacc = 0.0;

    // From the K-major image of a tile (16-byte chunk c of row r at c ^ (r & 7)): the MN-major image (32-byte chunk C
    // at C ^ (r & 3), its two halves in order) and, in the three-product mode, the remainder of every value in both
    // layouts.  The caller has made sure the GEMM1 that read this MN-major slot last has retired.
    Pipe pd;
    int nd = 0;  // tiles derived so far
    auto derive = [&]() {
      mbar_wait(&fullK[pd.st], pd.ph);
      const uint32_t hi_k = smem_u + pd.st * slotb, hi_mn = smem_u + mn_off + (mn2 ? (uint32_t)(nd & 1) * slotb : 0u);
      for (uint32_t off = et * 16; off < imgb; off += 512 * 16) {
        const float4 v = lds128(hi_k + off);
        const uint32_t rr = off >> 7, cc = ((off >> 4) & 7) ^ (rr & 7);
        const uint32_t pos = (((cc >> 1) ^ (rr & 3)) << 1) | (cc & 1);
        const uint32_t dst = (rr << 7) + (pos << 4);
        sts128(hi_mn + dst, v);
        if (!x1) {
          float4 r;
          r.x = __uint_as_float(to_tf32(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u)));
          r.y = __uint_as_float(to_tf32(v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u)));
          r.z = __uint_as_float(to_tf32(v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u)));
          r.w = __uint_as_float(to_tf32(v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u)));
          sts128(hi_k + imgb + off, r);
          sts128(hi_mn + imgb + dst, r);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&mnready[mn2 ? (nd & 1) : 0]);
      ++nd;
      pd.advance(kslots);
    };
    auto flush = [&](int j) {
      mbar_wait(gfull, (uint32_t)j & 1);
      tc_fence_after();
      const uint32_t g = tmem_base + lane_addr + p.col_g + c * NH;
      uint32_t t[NH];
#pragma unroll
      for (int e0 = 0; e0 + 8 <= NH; e0 += 8)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(t[e0]), "=r"(t[e0 + 1]), "=r"(t[e0 + 2]), "=r"(t[e0 + 3]), "=r"(t[e0 + 4]), "=r"(t[e0 + 5]),
                       "=r"(t[e0 + 6]), "=r"(t[e0 + 7])
                     : "r"(g + e0)
                     : "memory");
      if (NH % 8 == 4)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(t[NH - 4]), "=r"(t[NH - 3]), "=r"(t[NH - 2]), "=r"(t[NH - 1])
                     : "r"(g + NH - 4)
                     : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(gdrained);  // G' is free again; the additions run under the next GEMM2
#pragma unroll
      for (int e = 0; e < NH; ++e) acc[e] += (double)__uint_as_float(t[e]);
    };

    derive();
    for (int i = 0; i < nloc; ++i) {
      // two MN-major slots: the one of tile i + 1 was freed by GEMM1 of tile i - 1, which the previous turn waited for
      if (mn2 && i + 1 < nloc) derive();
      // b -> f, first half: f_hi replaces D1 (GEMM1 of this tile has retired)
      const int b = i & 1;
      mbar_wait(&d1full[b], ((uint32_t)i >> 1) & 1);
      tc_fence_after();
      const uint32_t col = b * V2_TS + c * 16;
      uint32_t fl[16];
      {
        float v[16];
        tmem_ld16(tmem_base + lane_addr + col, v);
        float w2 = 0.f;
        uint32_t fh[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float b2 = v[e] * v[e];
          w2 += b2;
          const float f = b2 * v[e];
          fh[e] = x1 ? to_tf32(f) : (__float_as_uint(f) & 0xffffe000u);
          fl[e] = to_tf32(f - __uint_as_float(fh[e]));
        }
        tmem_st16(tmem_base + lane_addr + col, fh);
        wacc += (double)w2;
      }
      // one MN-major slot: the images of the next tile now, while GEMM2 of the previous one runs
      if (!mn2 && i + 1 < nloc) derive();
      // second half: f_lo, which GEMM2 of the previous tile reads until it retires
      if (!x1) {
        if (i >= 1) {
          mbar_wait(gfull, (uint32_t)(i - 1) & 1);
          tc_fence_after();
        }
        tmem_st16(tmem_base + lane_addr + p.col_flo + c * 16, fl);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&fready[b]);
      if (i >= 1) flush(i - 1);
    }
    flush(nloc - 1);

    double* gp = p.gpart + ((int64_t)blockIdx.x * 128 + row) * (4 * NH) + c * NH;
#pragma unroll
    for (int e = 0; e < NH; e += 2) *reinterpret_cast<double2*>(gp + e) = make_double2(acc[e], acc[e + 1]);
    p.wpart[(int64_t)blockIdx.x * 512 + c * 128 + row] = wacc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// the packed kernel applies: TMEM columns and shared memory (three K-major slots in the three-product mode)
static bool vt2_shape(int64_t m, int products, int* kslots_out, int* mnslots_out = nullptr) {
  const int ng = (int)(m <= 32 ? 32 : m <= 64 ? 64 : m <= 96 ? 96 : m <= 112 ? 112 : 128);
  const int k1 = (int)round_up(m, 8), nb = (int)lpad(m);
  const int x1 = products == 1 ? 1 : 0;
  const int cols = x1 ? 2 * V2_TS + ng + k1 : 3 * V2_TS + ng + 2 * k1;
  const int slotb = (x1 ? 1 : 2) * 2 * nb * 128;
  const int budget = 227 * 1024 - 1024 /*alignment*/ - 512 /*barriers*/;
  const int total = budget / slotb;  // three K-major slots and one MN-major at least; a second MN-major one, then a
  const int mnslots = total >= 5 ? 2 : 1;  // fourth K-major one where there is room
  int kslots = total - mnslots;
  if (kslots > V2_MAX_SLOTS) kslots = V2_MAX_SLOTS;
  if (kslots_out) *kslots_out = kslots;
  if (mnslots_out) *mnslots_out = mnslots;
  return cols <= 512 && kslots >= 3;
}

int64_t varimax_pack_bytes(int64_t S, int64_t m) {
  if (m < 2 || m > 128 || !vt2_shape(m, 1, nullptr)) return 0;
  return ceil_div(S, V2_TS) * 2 * lpad(m) * 128;
}

int varimax_pack(const float* L, int64_t S, int64_t m, int64_t ld, float* packed, cudaStream_t stream) {
  const int64_t ntiles = ceil_div(S, V2_TS);
  const int nb = (int)lpad(m);
  const int64_t total = ntiles * nb * 16;
  vt_pack_kernel<<<(unsigned)imin(ceil_div(total, 256), 16 * (int64_t)num_sms()), 256, 0, stream>>>(L, S, ld, nb, ntiles, packed);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

static int vt_n2(int64_t m) {  // instantiated accumulator widths
  return m <= 32 ? 32 : m <= 64 ? 64 : m <= 96 ? 96 : m <= 112 ? 112 : 128;
}

int64_t varimax_tc_workspace_bytes(int64_t S, int64_t m) {
  (void)S;
  const int64_t ng = vt_n2(m);
  return 65536 /*R^T hi*/ + 65536 /*R^T lo*/ + (int64_t)num_sms() * (128 * ng + 512) * 8 + 1024;
}

bool varimax_tc_supported(const float* L, int64_t S, int64_t m, int64_t ld) {
  return m >= 2 && m <= 128 && S >= 1 && ld % 4 == 0 && ((uintptr_t)L % 16 == 0) && tensor_maps_available() &&
         S + 64 < (int64_t)1 << 31;
}

int varimax_sweep_tc(const float* L, const float* packed, int64_t S, int64_t m, int64_t ld, const double* R, double* Gout,
                     double* Wout, int accumulate, int products, void* workspace, int64_t workspace_bytes,
                     cudaStream_t stream) {
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
  // tiles of 64 features from the packed copy, where the caller made one and the shape fits
  int kslots = 0, mnslots = 1;
  if (packed && vt2_shape(m, products, &kslots, &mnslots) && env_int("XEOFS_VT_PAIR", 1)) {
    const int x1 = products == 1 ? 1 : 0;
    Vt2Params q{};
    q.nb = nb; q.ng = ng; q.k1 = k1; q.kslots = kslots; q.mnslots = mnslots; q.x1 = x1;
    q.pf = env_int("XEOFS_VT_PF", 2);
    q.ntiles = (int)ceil_div(S, V2_TS);
    q.col_flo = 2 * V2_TS;
    q.col_g = x1 ? 2 * V2_TS : 3 * V2_TS;
    q.col_rlo = q.col_g + ng;
    q.col_rhi = x1 ? q.col_g + ng : q.col_rlo + k1;
    const int grid2 = (int)imin(num_sms(), q.ntiles);
    double* gpart2 = (double*)ws;
    q.packed = packed; q.rhi = rhi; q.rlo = rlo; q.gpart = gpart2; q.wpart = gpart2 + (int64_t)grid2 * 128 * ng;
    vt_prep_R_kernel<<<16, 256, 0, stream>>>(R, (int)m, rhi, rlo, x1);
    XB_LAUNCH_CHECK();
    const size_t smem2 = (size_t)(kslots + mnslots) * (x1 ? 1 : 2) * 2 * nb * 128 + 1024 + 512;
#define XB_VT2(NHV)                                                                                                    \
  XB_CUDA(cudaFuncSetAttribute(varimax_tc2_kernel<NHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));     \
  varimax_tc2_kernel<NHV><<<grid2, V2_THREADS, smem2, stream>>>(q)
    switch (ng) {
      case 32: XB_VT2(8); break;
      case 64: XB_VT2(16); break;
      case 96: XB_VT2(24); break;
      case 112: XB_VT2(28); break;
      default: XB_VT2(32); break;
    }
#undef XB_VT2
    XB_LAUNCH_CHECK();
    vt_reduce_kernel<<<(unsigned)ceil_div((m * m + m + 3) / 4 * 32, 256), 256, 0, stream>>>(gpart2, q.wpart, grid2, (int)m, ng, Gout, Wout,
                                                                         accumulate, 4);
    XB_LAUNCH_CHECK();
    return XEOFS_OK;
  }
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
  vt_reduce_kernel<<<(unsigned)ceil_div((m * m + m + 3) / 4 * 32, 256), 256, 0, stream>>>(gpart, wpart, grid, (int)m, ng, Gout, Wout, accumulate, 2);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

}  // namespace xb
