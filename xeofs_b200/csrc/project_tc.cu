// tcgen05 (kind::tf32) versions of the two streaming products of the randomized range finder
// (sklearn.utils.extmath.randomized_range_finder's  M @ Q / M.T @ Q, reference call site
// xeofs/linalg/decomposer.py:141-146) with Scaler.transform (preprocessing/scaler.py:146-153) and the
// Sanitizer's NaN handling (preprocessing/sanitizer.py:124) folded into the operand path.
//
//   project_S:  Yt[j,s] = dscale[s] * sum_t (X[t,s]-pivot[s]) W[t,j]  + ccorr[s] * sum_t W[t,j]
//               D[M = 128 s][N = lp j],  K = t.   grid = ceil(S/128) CTAs, each sweeps all of T.
//   project_T:  Z[t,j]  = sum_s (X[t,s]-pivot[s]) dscale[s] Yt[j,s]   + sum_s ccorr[s] Yt[j,s]
//               D[M = 128 t][N = lp j],  K = s.   grid = ceil(T/128) x splits; the partial sums of the splits
//               are written to the workspace and added by a second (deterministic) kernel.
//
// One CTA = 10 warps:
//   warp 0      TMA producer: per K-chunk of 32 one box of raw X (16 KB) and one box of the small operand
//               (lp x 32, K-major, 128-byte swizzle) into a ring of shared-memory stages, mbarrier-signalled;
//   warps 2-9   operand stage: shared memory -> registers, subtract the pivot, NaN -> 0, (scale), round to
//               TF32 (and keep the fp32 remainder for the 3xTF32 mode) -> tcgen05.st into the TMEM A-operand
//               slot of the stage (lane = row of D, column = k) — the big operand never goes back to shared
//               memory; afterwards the same warps run the epilogue (tcgen05.ld, scale / rank-1 term, store);
//   warp 1      one thread issues tcgen05.mma.kind::tf32 with A from TMEM and B from shared memory, 4 (x1) or
//               12 (x3: hi*hi + hi*lo + lo*hi) instructions per stage, accumulator D in TMEM; tcgen05.commit
//               releases the stage.
// The tensor core adds into its fp32 accumulator with truncation, a bias of about 2e-8 per K=8 step (measured:
// 2.2e-5 after 8760 rows).  The 3xTF32 kernels therefore alternate between two TMEM accumulators and move each
// finished group of TC_FLUSH stages (64 K=8 steps) into fp32 registers of the epilogue warps (round-to-nearest adds).
#include <cuda.h>

#include "common.cuh"

namespace xb {

constexpr int TC_THREADS = 320;
constexpr int TC_KC = 32;        // K elements per stage (= one 128-byte swizzle atom of fp32)
constexpr int TC_TILE = 128;     // rows of D per CTA (TMEM lanes)
constexpr int TC_XBYTES = TC_TILE * TC_KC * 4;  // 16 KB of X per stage
constexpr int TC_MAX_STAGES = 6;
constexpr int TC_D_COLS = 128;   // TMEM columns per accumulator buffer (lp <= 128)
constexpr int TC_FLUSH = 16;     // 3xTF32: K-chunks accumulated in TMEM before the sum moves to fp32 registers

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

constexpr uint64_t HINT_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t HINT_EVICT_LAST = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem], kind::tf32, one CTA
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// K-major, 128-byte-swizzled operand tile (rows of 32 fp32 = 128 B, 8-row groups 1024 B apart):
// start address >> 4 | LBO (unused with swizzle) = 1 | SBO = 1024 >> 4 | descriptor version 1 | SWIZZLE_128B
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D fp32, A/B tf32, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_TILE >> 4) << 24);
}

struct TcParams {
  int64_t T, S;
  int lp;             // N of the MMA (multiple of 16, <= 128)
  int stages;
  int nchunks_total;  // K-chunks over the whole K extent
  int chunks_per_cta; // project_T: K-chunks per split (project_S: = nchunks_total)
  const float* pivot;   // project_S: [S];  project_T: zero-padded copy, multiple of 32 long
  const float* dscale;  // same
  const float* ccorr;   // project_S epilogue (may be null)
  const float* wsum;    // project_S epilogue: column sums of W [lp]
  float* out;           // project_S: Yt (ldo = ldy);  project_T: partial sums [split][tiles*128][lp]
  int64_t ldo;
  uint32_t tmem_cols;
};

// ------------------------------------------------------------------------------------------------ the kernel
template <int NS, bool SIDE_T>
__global__ void __launch_bounds__(TC_THREADS, NS == 1 ? 2 : 1)
project_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapBhi,
                  const __grid_constant__ CUtensorMap mapBlo, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stages = p.stages, lp = p.lp;
  const int bbytes = lp * TC_KC * 4;
  uint8_t* xs = smem;                                   // [stages][16 KB]
  uint8_t* bhi = xs + (size_t)stages * TC_XBYTES;       // [stages][lp*128 B]
  uint8_t* blo = bhi + (size_t)stages * bbytes;         // [stages][lp*128 B]   (NS == 3)
  uint8_t* pd = blo + (NS == 3 ? (size_t)stages * bbytes : 0);  // [stages][256 B] pivot | dscale (SIDE_T)
  uint64_t* bars = (uint64_t*)(pd + (size_t)stages * 256);
  uint64_t* full = bars;                       // TMA bytes landed
  uint64_t* empty = bars + TC_MAX_STAGES;      // MMAs of the stage retired
  uint64_t* aready = bars + 2 * TC_MAX_STAGES; // A operand of the stage is in TMEM
  uint64_t* dfull = bars + 3 * TC_MAX_STAGES;  // [2] accumulator buffer complete
  uint64_t* dempty = dfull + 2;                // [2] accumulator buffer drained into registers (NS == 3)
  uint32_t* tmem_slot = (uint32_t*)(dempty + 2);

  const int64_t tile0 = (int64_t)blockIdx.x * TC_TILE;  // first row of D: s (project_S) or t (project_T)
  const int chunk0 = SIDE_T ? blockIdx.y * p.chunks_per_cta : 0;
  const int nchunks = SIDE_T ? min(p.chunks_per_cta, p.nchunks_total - chunk0) : p.nchunks_total;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&aready[i], 8);
    }
    mbar_init(&dfull[0], 1);
    mbar_init(&dfull[1], 1);
    mbar_init(&dempty[0], 8);
    mbar_init(&dempty[1], 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t a_cols_per_stage = TC_KC * (NS == 3 ? 2 : 1);
  const uint32_t a_col0 = NS == 3 ? 2 * TC_D_COLS : TC_D_COLS;  // first TMEM column of the A-operand ring

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      const uint32_t tx = TC_XBYTES + bbytes * (NS == 3 ? 2 : 1) + (SIDE_T ? 256 : 0);
      for (int c = 0; c < nchunks; ++c) {
        const int st = c % stages;
        const uint32_t ph = (c / stages) & 1;
        mbar_wait(&empty[st], ph ^ 1);
        mbar_expect_tx(&full[st], tx);
        const int k0 = (chunk0 + c) * TC_KC;
        if (!SIDE_T) {
          tma_load_2d(xs + (size_t)st * TC_XBYTES, &mapX, (int)tile0, k0, &full[st], HINT_EVICT_FIRST);
        } else {
          tma_load_2d(xs + (size_t)st * TC_XBYTES, &mapX, k0, (int)tile0, &full[st], HINT_EVICT_FIRST);
          bulk_load_1d(pd + st * 256, p.pivot + k0, 128, &full[st]);
          bulk_load_1d(pd + st * 256 + 128, p.dscale + k0, 128, &full[st]);
        }
        tma_load_2d(bhi + (size_t)st * bbytes, &mapBhi, k0, 0, &full[st], HINT_EVICT_LAST);
        if (NS == 3) tma_load_2d(blo + (size_t)st * bbytes, &mapBlo, k0, 0, &full[st], HINT_EVICT_LAST);
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(lp);
      for (int c = 0; c < nchunks; ++c) {
        const int st = c % stages;
        const uint32_t ph = (c / stages) & 1;
        uint32_t d_tmem = tmem_base;
        bool first = c == 0;
        if (NS == 3) {
          const int g = c / TC_FLUSH, buf = g & 1;
          first = (c % TC_FLUSH) == 0;
          if (first) {
            mbar_wait(&dempty[buf], (((uint32_t)g >> 1) & 1) ^ 1);  // registers hold what this buffer had
            tc_fence_after();
          }
          d_tmem = tmem_base + buf * TC_D_COLS;
        }
        mbar_wait(&full[st], ph);
        mbar_wait(&aready[st], ph);
        tc_fence_after();
        const uint32_t a_hi = tmem_base + a_col0 + st * a_cols_per_stage;
        const uint64_t dh = make_b_desc(smem_u32(bhi + (size_t)st * bbytes));
        const uint64_t dl = NS == 3 ? make_b_desc(smem_u32(blo + (size_t)st * bbytes)) : 0;
#pragma unroll
        for (int k = 0; k < TC_KC / 8; ++k) {
          // +32 bytes (8 tf32) along K inside the swizzle atom = +2 in the (address >> 4) field
          mma_tf32_ts(d_tmem, a_hi + k * 8, dh + 2 * k, idesc, !(first && k == 0));
          if (NS == 3) {
            mma_tf32_ts(d_tmem, a_hi + k * 8, dl + 2 * k, idesc, 1);
            mma_tf32_ts(d_tmem, a_hi + TC_KC + k * 8, dh + 2 * k, idesc, 1);
          }
        }
        mma_commit(&empty[st]);
        if (NS == 3 && ((c % TC_FLUSH) == TC_FLUSH - 1 || c == nchunks - 1)) mma_commit(&dfull[(c / TC_FLUSH) & 1]);
      }
      if (NS == 1) mma_commit(&dfull[0]);
    }
  } else {
    // ===================================================================== operand stage + epilogue
    const int q = warp & 3;             // TMEM lane quarter this warp may touch
    const int half = (warp - 2) >> 2;   // which 16 of the 32 K values of a stage
    const int row = q * 32 + lane;      // row of D / TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float piv = 0.f;
    if (!SIDE_T) piv = (tile0 + row < p.S) ? p.pivot[tile0 + row] : 0.f;

    const int groups = lp >> 4;  // groups of 8 accumulator columns owned by this warp (its half of lp)
    float acc[NS == 3 ? 64 : 1];
#pragma unroll
    for (int i = 0; i < (NS == 3 ? 64 : 1); ++i) acc[i] = 0.f;
    const int nflush = (nchunks + TC_FLUSH - 1) / TC_FLUSH;
    int next_flush = 0;
    // registers += accumulator buffer of flush group g (after the MMA warp committed it), then hand the buffer back
    auto flush = [&](int g) {
      const int buf = g & 1;
      mbar_wait(&dfull[buf], ((uint32_t)g >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int gi = 0; gi < 8; ++gi) {
        if (gi < groups) {
          float v[8];
          tmem_ld8(tmem_base + lane_addr + buf * TC_D_COLS + half * (lp >> 1) + gi * 8, v);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[(NS == 3 ? gi * 8 + e : 0)] += v[e];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dempty[buf]);
    };

    for (int c = 0; c < nchunks; ++c) {
      const int st = c % stages;
      const uint32_t ph = (c / stages) & 1;
      if (NS == 3 && next_flush < nflush && c >= (next_flush + 1) * TC_FLUSH + 1) flush(next_flush++);
      mbar_wait(&full[st], ph);
      tc_fence_after();
      uint32_t hi[16], lo[16];
      if (!SIDE_T) {
        // X stage = [32 t][128 s] fp32; this thread owns column `row`, rows half*16 .. +15
        const float* src = (const float*)(xs + (size_t)st * TC_XBYTES) + (half * 16) * TC_TILE + row;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          float v = src[r * TC_TILE] - piv;
          v = (v == v) ? v : 0.f;
          hi[r] = to_tf32(v);
          if (NS == 3) lo[r] = __float_as_uint(v - __uint_as_float(hi[r]));
        }
      } else {
        // X stage = [128 t][32 s] fp32, 16-byte chunks XOR-swizzled by (row & 7); this thread owns row `row`,
        // chunks half*4 .. +3
        const float* src = (const float*)(xs + (size_t)st * TC_XBYTES) + row * TC_KC;
        const float* pv = (const float*)(pd + st * 256);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int ch = half * 4 + cc;
          const float4 x = *reinterpret_cast<const float4*>(src + ((ch ^ (row & 7)) << 2));
          const float4 pq = *reinterpret_cast<const float4*>(pv + ch * 4);
          const float4 dq = *reinterpret_cast<const float4*>(pv + 32 + ch * 4);
          const float xa[4] = {x.x, x.y, x.z, x.w}, pa[4] = {pq.x, pq.y, pq.z, pq.w}, da[4] = {dq.x, dq.y, dq.z, dq.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float v = xa[e] - pa[e];
            v = (v == v) ? v * da[e] : 0.f;
            hi[cc * 4 + e] = to_tf32(v);
            if (NS == 3) lo[cc * 4 + e] = __float_as_uint(v - __uint_as_float(hi[cc * 4 + e]));
          }
        }
      }
      const uint32_t a_hi = tmem_base + lane_addr + a_col0 + st * a_cols_per_stage + half * 16;
      tmem_st16(a_hi, hi);
      if (NS == 3) tmem_st16(a_hi + TC_KC, lo);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&aready[st]);
    }

    // ---- epilogue: D (128 x lp fp32: TMEM for x1, registers for x3) -> global
    if (NS == 3) {
      while (next_flush < nflush) flush(next_flush++);
    } else {
      mbar_wait(&dfull[0], 0);
      tc_fence_after();
    }
    const int64_t rrow = tile0 + row;  // s (project_S) or t (project_T)
    const bool ok = SIDE_T ? true : rrow < p.S;
    float ds = 0.f, cs = 0.f;
    if (!SIDE_T && ok) {
      ds = p.dscale[rrow];
      cs = p.ccorr ? p.ccorr[rrow] : 0.f;
    }
    float* dstT = SIDE_T ? p.out + ((int64_t)blockIdx.y * gridDim.x * TC_TILE + rrow) * lp : nullptr;
#pragma unroll
    for (int gi = 0; gi < 8; ++gi) {
      if (gi < groups) {
        const int j0 = half * (lp >> 1) + gi * 8;
        float v[8];
        if (NS == 3) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = acc[(NS == 3 ? gi * 8 + e : 0)];
        } else {
          tmem_ld8(tmem_base + lane_addr + j0, v);
        }
        if (!SIDE_T) {
          if (ok) {
#pragma unroll
            for (int e = 0; e < 8; ++e) p.out[(int64_t)(j0 + e) * p.ldo + rrow] = fmaf(ds, v[e], cs * p.wsum[j0 + e]);
          }
        } else {
          *reinterpret_cast<float4*>(dstT + j0) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(dstT + j0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ small helpers
// W (T x ldw, time-side) -> Wt_hi / Wt_lo (lp x Tpad, K-major for the MMA: t contiguous), TF32-rounded value and
// fp32 remainder; pad columns (t >= T) and pad rows (j >= l... already zero in W) are zero.  wsum[j] = sum_t W[t,j].
__global__ void __launch_bounds__(256)
prep_W_kernel(const float* __restrict__ W, int64_t T, int64_t ldw, int lp, int64_t Tpad, float* __restrict__ Whi,
              float* __restrict__ Wlo) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t t0 = (int64_t)blockIdx.x * 32;
  const int j0 = blockIdx.y * 32;
  for (int r = ty; r < 32; r += 8) {
    const int64_t t = t0 + r;
    const int j = j0 + tx;
    tile[r][tx] = (t < T && j < lp) ? W[t * ldw + j] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int j = j0 + r;
    const int64_t t = t0 + tx;
    if (j < lp && t < Tpad) {
      const float v = tile[tx][r];
      const float h = __uint_as_float(to_tf32(v));
      Whi[(int64_t)j * Tpad + t] = h;
      if (Wlo) Wlo[(int64_t)j * Tpad + t] = v - h;
    }
  }
}

// zero-padded copies of pivot / dscale (length Spad, multiple of 32) for the 128-byte bulk copies of project_T
__global__ void pad_vectors_kernel(const float* __restrict__ pivot, const float* __restrict__ dscale, int64_t S,
                                   int64_t Spad, float* __restrict__ ppad, float* __restrict__ dpad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Spad) {
    ppad[i] = i < S ? pivot[i] : 0.f;
    dpad[i] = i < S ? dscale[i] : 0.f;
  }
}

// Yt (lp x ldy) -> TF32-rounded copy and fp32 remainder (lp x Spad), zero beyond S
__global__ void split_Y_kernel(const float* __restrict__ Yt, int64_t S, int64_t ldy, int64_t Spad, float* __restrict__ Yhi,
                               float* __restrict__ Ylo) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int64_t j = blockIdx.y;
  if (i >= Spad) return;
  float v[4], h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[e] = (i + e < S) ? Yt[j * ldy + i + e] : 0.f;
    h[e] = __uint_as_float(to_tf32(v[e]));
    l[e] = v[e] - h[e];
  }
  *reinterpret_cast<float4*>(Yhi + j * Spad + i) = make_float4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<float4*>(Ylo + j * Spad + i) = make_float4(l[0], l[1], l[2], l[3]);
}

// Z[t, j] = sum_split P[split][t][j] + r[j]   (r only on the valid samples)
__global__ void reduce_partials_kernel(const float* __restrict__ P, int splits, int64_t rows_pad, int lp, int64_t T,
                                       const float* __restrict__ r, const uint8_t* __restrict__ row_valid,
                                       float* __restrict__ Z, int64_t ldz) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * lp) return;
  const int64_t t = i / lp;
  const int j = (int)(i % lp);
  float acc = (r && (!row_valid || row_valid[t])) ? r[j] : 0.f;
  for (int s = 0; s < splits; ++s) acc += P[((int64_t)s * rows_pad + t) * lp + j];
  Z[t * ldz + j] = acc;
}

// project_simt.cu
int launch_colsum(const float* W, int64_t T, int64_t ldw, int lp, const uint8_t* row_valid, float* out, cudaStream_t stream);
int launch_ccorr_dot(const float* Yt, int64_t S, int64_t ldy, const float* ccorr, int lp, float* out, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 tensor map: inner (contiguous) extent `inner`, `outer` rows `ld` elements apart
static int make_map(CUtensorMap* m, const float* base, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                    int box_outer, bool swizzle128) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return XEOFS_E_UNSUPPORTED;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for extent %lld x %lld ld %lld box %d x %d", (int)r, (long long)inner,
              (long long)outer, (long long)ld, box_inner, box_outer);
    return XEOFS_E_CUDA;
  }
  return XEOFS_OK;
}

bool tc_supported(int64_t T, int64_t S, int64_t ldx, const float* X, int64_t l) {
  (void)T; (void)S;
  return l >= 1 && l <= 128 && ldx % 4 == 0 && ((uintptr_t)X % 16 == 0) && get_encode() != nullptr;
}

static inline int64_t align256(int64_t b) { return round_up(b, 256); }

struct TGeom {
  int64_t t_tiles, rows_pad, Spad;
  int chunks_total, chunks_per_cta, splits;
};
static TGeom t_geometry(int64_t T, int64_t S, bool x3) {
  TGeom g;
  g.t_tiles = ceil_div(T, TC_TILE);
  g.rows_pad = g.t_tiles * TC_TILE;
  g.Spad = round_up(S, TC_KC);
  g.chunks_total = (int)(g.Spad / TC_KC);
  int64_t want = ceil_div(4 * (int64_t)num_sms(), g.t_tiles);
  if (want < 1) want = 1;
  if (want > g.chunks_total) want = g.chunks_total;
  g.chunks_per_cta = (int)ceil_div(g.chunks_total, want);
  // single-TF32 kernels keep one TMEM accumulator for the whole K range of a CTA: bound its truncation bias
  if (!x3 && g.chunks_per_cta > 1024) g.chunks_per_cta = 1024;
  g.splits = (int)ceil_div(g.chunks_total, g.chunks_per_cta);
  return g;
}

int64_t tc_workspace_bytes(int64_t T, int64_t S, int64_t l, int algo) {
  const int64_t lp = lpad(l);
  const bool x3 = (algo == XEOFS_ALGO_TF32X3 || algo == XEOFS_ALGO_AUTO);
  const int64_t Tpad = round_up(T, TC_KC);
  // project_S: wsum | Wt_hi | Wt_lo
  const int64_t bs = align256(lp * 4) + (x3 ? 2 : 1) * align256(lp * Tpad * 4);
  // project_T: rvec | pivot_pad | dscale_pad | partials | Yhi | Ylo
  const TGeom g = t_geometry(T, S, false);  // the finer split needs the larger partial-sum buffer
  int64_t bt = align256(lp * 4) + 2 * align256(g.Spad * 4) + align256((int64_t)g.splits * g.rows_pad * lp * 4);
  if (x3) bt += 2 * align256(lp * g.Spad * 4);
  return (bs > bt ? bs : bt) + 256;
}

static int pick_stages(int lp, int ns, bool side_t, size_t* smem_bytes) {
  const int per_stage = TC_XBYTES + lp * TC_KC * 4 * (ns == 3 ? 2 : 1) + 256;
  // x1: two CTAs per SM (TMEM 2 x 256 columns) -> about 110 KB each; x3: one CTA per SM
  const int budget = ns == 1 ? 110 * 1024 : 200 * 1024;
  const int max_by_tmem = 4;  // x1: (256 - 128) / 32;  x3: (512 - 2 * 128) / 64
  int st = (budget - 2048) / per_stage;
  if (st > max_by_tmem) st = max_by_tmem;
  if (st > TC_MAX_STAGES) st = TC_MAX_STAGES;
  if (st < 2) st = 2;
  (void)side_t;
  *smem_bytes = (size_t)st * per_stage + 1024 /*alignment*/ + 256 /*barriers*/;
  return st;
}

template <int NS, bool SIDE_T>
static int launch_tc(const CUtensorMap& mx, const CUtensorMap& mh, const CUtensorMap& ml, const TcParams& p, dim3 grid,
                     size_t smem, cudaStream_t stream) {
  XB_CUDA(cudaFuncSetAttribute(project_tc_kernel<NS, SIDE_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  project_tc_kernel<NS, SIDE_T><<<grid, TC_THREADS, smem, stream>>>(mx, mh, ml, p);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

int project_S_tc(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot, const float* dscale,
                 const float* ccorr, const uint8_t* row_valid, const float* W, int64_t ldw, int64_t l, float* Yt,
                 int64_t ldy, void* workspace,
                 int64_t workspace_bytes, int algo, cudaStream_t stream) {
  (void)workspace_bytes;
  const int lp = (int)lpad(l);
  const int ns = algo == XEOFS_ALGO_TF32X3 ? 3 : 1;
  const int64_t Tpad = round_up(T, TC_KC);
  uint8_t* ws = (uint8_t*)workspace;
  float* wsum = (float*)ws;
  float* Whi = (float*)(ws + align256(lp * 4));
  float* Wlo = ns == 3 ? (float*)((uint8_t*)Whi + align256(lp * Tpad * 4)) : nullptr;
  int rc = launch_colsum(W, T, ldw, lp, row_valid, wsum, stream);
  if (rc) return rc;
  prep_W_kernel<<<dim3((unsigned)(Tpad / 32), (unsigned)ceil_div(lp, 32)), 256, 0, stream>>>(W, T, ldw, lp, Tpad, Whi, Wlo);
  XB_LAUNCH_CHECK();
  CUtensorMap mx, mh, ml;
  rc = make_map(&mx, X, S, T, ldx, TC_TILE, TC_KC, false);
  if (rc) return rc;
  rc = make_map(&mh, Whi, Tpad, lp, Tpad, TC_KC, lp, true);
  if (rc) return rc;
  ml = mh;
  if (ns == 3) {
    rc = make_map(&ml, Wlo, Tpad, lp, Tpad, TC_KC, lp, true);
    if (rc) return rc;
  }
  TcParams p{};
  p.T = T; p.S = S; p.lp = lp;
  size_t smem;
  p.stages = pick_stages(lp, ns, false, &smem);
  p.nchunks_total = (int)(Tpad / TC_KC);
  p.chunks_per_cta = p.nchunks_total;
  p.pivot = pivot; p.dscale = dscale; p.ccorr = ccorr; p.wsum = wsum;
  p.out = Yt; p.ldo = ldy;
  p.tmem_cols = ns == 1 ? 256 : 512;
  dim3 grid((unsigned)ceil_div(S, TC_TILE));
  return ns == 3 ? launch_tc<3, false>(mx, mh, ml, p, grid, smem, stream) : launch_tc<1, false>(mx, mh, ml, p, grid, smem, stream);
}

int project_T_tc(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot, const float* dscale,
                 const float* ccorr, const uint8_t* row_valid, const float* Yt, int64_t ldy, int64_t l, float* Z,
                 int64_t ldz, void* workspace,
                 int64_t workspace_bytes, int algo, cudaStream_t stream) {
  (void)workspace_bytes;
  const int lp = (int)lpad(l);
  int ns = algo == XEOFS_ALGO_TF32X3 ? 3 : 1;
  const TGeom g = t_geometry(T, S, ns == 3);
  XB_CHECK_ARG(g.splits <= 65535, "project_T: too many splits");
  uint8_t* ws = (uint8_t*)workspace;
  float* rvec = (float*)ws; ws += align256(lp * 4);
  float* ppad = (float*)ws; ws += align256(g.Spad * 4);
  float* dpad = (float*)ws; ws += align256(g.Spad * 4);
  float* part = (float*)ws; ws += align256((int64_t)g.splits * g.rows_pad * lp * 4);
  float* Yhi = (float*)ws; ws += align256(lp * g.Spad * 4);
  float* Ylo = (float*)ws;
  pad_vectors_kernel<<<(unsigned)ceil_div(g.Spad, 256), 256, 0, stream>>>(pivot, dscale, S, g.Spad, ppad, dpad);
  XB_LAUNCH_CHECK();
  // the small operand straight from the caller's buffer when it can be a TMA source and one product is enough
  const bool direct = ns == 1 && ldy % 4 == 0 && ((uintptr_t)Yt % 16 == 0);
  CUtensorMap mx, mh, ml;
  int rc = make_map(&mx, X, S, T, ldx, TC_KC, TC_TILE, true);
  if (rc) return rc;
  if (direct) {
    rc = make_map(&mh, Yt, S, lp, ldy, TC_KC, lp, true);
    if (rc) return rc;
    ml = mh;
  } else {
    split_Y_kernel<<<dim3((unsigned)ceil_div(g.Spad / 4, 256), (unsigned)lp), 256, 0, stream>>>(Yt, S, ldy, g.Spad, Yhi, Ylo);
    XB_LAUNCH_CHECK();
    rc = make_map(&mh, Yhi, g.Spad, lp, g.Spad, TC_KC, lp, true);
    if (rc) return rc;
    rc = make_map(&ml, Ylo, g.Spad, lp, g.Spad, TC_KC, lp, true);
    if (rc) return rc;
  }
  if (ccorr) {
    rc = launch_ccorr_dot(Yt, S, ldy, ccorr, lp, rvec, stream);
    if (rc) return rc;
  }
  TcParams p{};
  p.T = T; p.S = S; p.lp = lp;
  size_t smem;
  p.stages = pick_stages(lp, ns, true, &smem);
  p.nchunks_total = g.chunks_total;
  p.chunks_per_cta = g.chunks_per_cta;
  p.pivot = ppad; p.dscale = dpad; p.ccorr = nullptr; p.wsum = nullptr;
  p.out = part; p.ldo = lp;
  p.tmem_cols = ns == 1 ? 256 : 512;
  dim3 grid((unsigned)g.t_tiles, (unsigned)g.splits);
  rc = ns == 3 ? launch_tc<3, true>(mx, mh, ml, p, grid, smem, stream) : launch_tc<1, true>(mx, mh, ml, p, grid, smem, stream);
  if (rc) return rc;
  reduce_partials_kernel<<<(unsigned)ceil_div(T * lp, 256), 256, 0, stream>>>(part, g.splits, g.rows_pad, lp, T,
                                                                              ccorr ? rvec : nullptr, row_valid, Z, ldz);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

}  // namespace xb
