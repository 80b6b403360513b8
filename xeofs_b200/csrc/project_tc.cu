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
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace xb {

// warps per TMEM lane quarter in the operand stage: the 3xTF32 kernels run one CTA per SM and need the extra warps to
// hide the latencies of their longer conversion chain
// (wide: one CTA per SM — every instantiation but the plain single-TF32 project_S, which runs two CTAs of 8 operand warps)
__host__ __device__ constexpr int tc_nw(bool wide) { return wide ? 4 : 2; }
__host__ __device__ constexpr int tc_threads(bool wide) { return 64 + 128 * tc_nw(wide); }
// (16 operand warps were tried for the single-TF32 project_T: 3-5 % slower than 8)
__host__ __device__ constexpr bool tc_wide(int ns, bool side_t, bool rn) { return ns >= 2; }
constexpr int TC_KC = 32;        // K elements per stage (= one 128-byte swizzle atom of fp32)
constexpr int TC_XBYTES = TC_TILE * TC_KC * 4;  // 16 KB of X per stage
constexpr int TC_MAX_STAGES = 6;
constexpr int TC_FLUSH = 16;     // 3xTF32: K-chunks accumulated in TMEM before the sum moves to fp32 registers
constexpr int TC_FLUSH_PK = 32;  // the same for the packed mode (one accumulator buffer: the issuer waits for the move)

struct TcParams {
  int64_t T, S;
  int lp;             // N of the MMA (multiple of 16, <= 128)
  int stages;
  int nchunks_total;  // stages' worth of K (32*KB values each) over the whole K extent
  int chunks_per_cta; // project_T: K-chunks per split (project_S: = nchunks_total)
  int dcols;          // TMEM columns per accumulator buffer (lp rounded up to 32)
  const float* pivot;   // project_S: [S];  project_T: zero-padded copy (multiple of 128 long)
  const float* dscale;  // same
  const float* ccorr;   // project_S epilogue (may be null)
  const float* wsum;    // project_S epilogue: column sums of W [lp]
  const uint8_t* chunk_flags;  // project_S: [K-chunk] 1 if the 32 samples hold one that is NaN throughout (null: none)
  // fused statistics (project_S_stats): Scaler.fit / Sanitizer masks / total variance from the same read of X
  const double* featw;   // [S] coslat * weights (null = 1)
  int stat_flags;        // XEOFS_F_CENTER | XEOFS_F_STANDARDIZE
  float* mean_out;       // [S]
  float* std_out;        // [S]
  uint8_t* valid_out;    // [S]
  float* pivot_out;      // [S] what later passes subtract (mean when centring, else the first sample)
  float* dscale_out;     // [S]
  float* ccorr_out;      // [S]
  double* scalars_out;   // [4] total variance, valid features, max / min count (pre-initialised)
  int32_t* row_delta;    // [T] NaN count of each sample minus that of the first sample (zero-initialised)
  int32_t* base_nan;     // [1] NaN count of the first sample (zero-initialised)
  float* out;           // project_S: Yt (ldo = ldy);  project_T: partial sums [split][tiles*128][lp]
  int64_t ldo;
  uint32_t tmem_cols;
  const float* X;       // project_T: rows are fetched with one bulk copy each
  int64_t ldx;
  const float* bimg_hi; // small operand as a sequence of shared-memory images, one per 32-wide K slab:
  const float* bimg_lo; //   [slab][lp rows][32 k], 16-byte chunks XOR-swizzled by (row & 7)  (lo: 3xTF32 only)
  int b_bulk;           // fetch the operand image of a stage with one bulk copy (else: tensor-map rows of 1 KB)
  int cl;               // CTAs per cluster (1, 2 or 4): they share the operand image through TMA multicast
  uint16_t* copy16;     // MODE_WCOPY: the fp16 copy this pass writes (T x ldc16)
  int64_t ldc16;
  const float* c0;      // statistics pass writing the copy: the power of two of every feature (from a pre-sample)
  float* ic16_out;      //   [S] what the copy has to be multiplied with: dscale / c0
  float* cc16_out;      //   [S] the rank-1 term of the copy: (first sample - mean) dscale
  const float* igs;     // MODE_H16: 1 / the power-of-two column scale of the small operand's fp16 image
};


// ------------------------------------------------------------------------------------------------ the kernel
// NS: 1 = single TF32 product, 3 = 3xTF32.  SIDE_T: false = project_S, true = project_T.
// KB: 32-wide K slabs per pipeline stage (project_T reads KB*128 contiguous bytes of every row per TMA box).
// RN (single TF32 only): round the operands to TF32 (to nearest) instead of letting the tensor core truncate them — the
// product is then unbiased (XEOFS_ALGO_TF32X1R: sums that are read as numbers, not only as a subspace).
// PK (3xTF32 only, lp <= 96): the two products that share the big operand's upper part, a_hi b_hi and a_hi b_lo, are
// ONE instruction of N = 2 lp against the operand image [b_hi rows | b_lo rows]; a_lo b_hi follows with N = lp into
// the first half of the accumulator.  tcgen05.mma has a floor of ~63 cycles per instruction whatever N is below ~64
// (profiles/r01_mma_issue_probe.txt): two instructions per K step instead of three.  The accumulator is then 2 lp
// columns wide, so there is one buffer (not two taking turns): the issuer waits while the epilogue warps move it into
// registers, every TC_FLUSH_PK slabs.
// TFAST (project_T): the caller vouches that the field holds no NaN (every feature and every sample valid), so the
// operand stage skips its per-value test.  The operand warps of project_T are issue-bound (measured: three extra
// integer instructions per value cost 40 % of a pass): every instruction taken out of their loop counts.  For the
// same reason dscale is folded into the small operand's image (tile_Y_kernel) unless that operand has to stay
// TF32-exact (NS == 2).
// RN = 2 (XEOFS_ALGO_TF32X1F): the field is a materialised, already TF32-rounded copy of the preprocessed matrix (pivot 0,
// dscale 1, no NaN): the operand stage only moves it from shared memory into TMEM; accumulators flushed as for RN = 1.
// MODE (single TF32 only) — the half-precision copy of the preprocessed matrix for the power iterations:
//   MODE_WCOPY  project_T, fp32 field: besides its product the pass writes A16[t,s] = fp16((X[t,s] - pivot[s]) e16[s]),
//               e16 = dscale x a power of two per feature that puts the column's standard deviation at 512 (values up
//               to 127 sigma representable, beyond that saturated); NaN -> 0.
//   MODE_H16    the field IS that copy (T x S fp16, features contiguous): the operand stage is a plain move of packed
//               pairs into TMEM, the products are kind::f16 (K = 16 per instruction, a slab = 64 K values = the same
//               128 bytes per row and 32 TMEM columns as 32 fp32 values), the small operand's image is fp16.
//               fp16 holds the 11 significant bits the tensor core reads of a TF32 operand — rounded to nearest here,
//               truncated there — so these passes lose nothing against the single-TF32 ones and move half the bytes.
constexpr int MODE_F32 = 0, MODE_WCOPY = 1, MODE_H16 = 2;

template <int NS, bool SIDE_T, int KB, bool STATS = false, int RN = 0, bool PK = false, bool TFAST = false, int MODE = MODE_F32>
__global__ void __launch_bounds__(tc_threads(tc_wide(NS, SIDE_T, RN != 0)), (NS == 1 && !SIDE_T && RN == 0) ? 2 : 1)
project_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapBhi,
                  const __grid_constant__ CUtensorMap mapBlo, const TcParams p) {
  static_assert(MODE == MODE_F32 || (NS == 1 && RN == 0 && !PK && (!STATS || MODE == MODE_WCOPY)),
                "the fp16 copy serves the single-product passes");
  static_assert(MODE != MODE_WCOPY || SIDE_T || STATS, "the copy is written by the statistics pass or the first project_T pass");
  constexpr bool H16 = MODE == MODE_H16;
  constexpr int KEL = H16 ? 2 * TC_KC : TC_KC;      // K values per slab (32 fp32 or 64 fp16: 128 bytes either way)
  static_assert(SIDE_T || KB == 1, "project_S stages are 32 rows of t");
  static_assert(!STATS || (NS == 1 && !SIDE_T), "the statistics ride on the first (single TF32) project_S pass");
  static_assert(!PK || NS == 3, "operand packing is the 3xTF32 mode's");
  constexpr int NBUF = PK ? 1 : 2;                  // accumulator buffers of the flushing modes
  constexpr int NPART = NS >= 2 ? 2 : 1;            // parts of the big operand kept per value (hi | lo)
  constexpr int BPART = NS == 3 ? 2 : 1;            // parts of the small operand (NS == 2: it is TF32-exact already)
  // FL: two TMEM accumulators that take turns, each moved into fp32 registers (round-to-nearest adds) after TC_FLUSH
  // slabs — the tensor core's own adds truncate.  All multi-product modes, and the rounded single product.
  constexpr bool FL = NS >= 2 || RN != 0;
  constexpr int NW = tc_nw(tc_wide(NS, SIDE_T, RN != 0));  // operand-stage warps per TMEM lane quarter
  constexpr int KW = TC_KC / NW;                    // K values of a slab converted by one warp
  // project_T: a stage holds 128 rows of KB*32 (+4) floats, KB*128 + 16 bytes apart: TMA fetches them as 128 long
  // pieces (the engine's cost is per piece, about 7 cycles, whatever its length), one thread then reads one row, and
  // the odd multiple of 16 bytes spreads the rows of a quarter-warp over all banks without a swizzle
  constexpr int XPITCH = SIDE_T ? KB * 128 + 16 : TC_TILE * 4;
  constexpr int XB = SIDE_T ? TC_TILE * XPITCH : TC_XBYTES;  // bytes of X per stage
  constexpr int ACOLS = TC_KC * KB * NPART;         // TMEM columns of one A-operand slot
  constexpr int FLUSH_STAGES = (PK ? TC_FLUSH_PK : TC_FLUSH) / KB;  // stages per accumulator flush group (NS >= 2)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stages = p.stages, lp = p.lp;
  const int bbytes = lp * TC_KC * 4;                    // one K-major slab of the small operand
  uint8_t* xs = smem;                                   // [stages][XB]
  uint8_t* bs = xs + (size_t)stages * XB;               // [stages][BPART][KB][bbytes]
  uint8_t* pd = bs + (size_t)stages * BPART * KB * bbytes;  // [stages][pivot KB*128 B | dscale KB*128 B] (SIDE_T)
  // MODE_WCOPY: two staging tiles [128 rows][KB*64 (+16) bytes] through which the fp16 values of a stage leave as whole
  // 128-byte row segments (a thread owns 32 bytes of a row: stored directly they are scattered 16-byte writes)
  constexpr int CPITCH = KB * 64 + 16;
  uint8_t* cst = pd + (SIDE_T ? (size_t)stages * KB * 256 : 0);
  uint64_t* bars = (uint64_t*)(cst + ((MODE == MODE_WCOPY && SIDE_T) ? 2 * TC_TILE * CPITCH : 0));
  uint64_t* full = bars;                       // TMA bytes landed
  uint64_t* empty = bars + TC_MAX_STAGES;      // MMAs of the stage retired
  uint64_t* aready = bars + 2 * TC_MAX_STAGES; // A operand of the stage is in TMEM
  uint64_t* dfull = bars + 3 * TC_MAX_STAGES;  // [2] accumulator buffer complete
  uint64_t* dempty = dfull + 2;                // [2] accumulator buffer drained into registers (NS == 3)
  uint32_t* tmem_slot = (uint32_t*)(dempty + 2);

  const int64_t tile0 = (int64_t)blockIdx.x * TC_TILE;  // first row of D: s (project_S) or t (project_T)
  const int chunk0 = SIDE_T ? blockIdx.y * p.chunks_per_cta : 0;
  const int nchunks = SIDE_T ? min(p.chunks_per_cta, p.nchunks_total - chunk0) : p.nchunks_total;

  const int cl = p.cl;  // cluster size; > 1: every CTA of the cluster fetches 1/cl of the operand image for all
  const uint16_t cl_mask = (uint16_t)((1u << cl) - 1u);
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], cl);  // released by the MMAs of every CTA that the stage's operand image was written to
      mbar_init(&aready[i], 4 * NW);
    }
    mbar_init(&dfull[0], 1);
    mbar_init(&dfull[1], 1);
    mbar_init(&dempty[0], 4 * NW);
    mbar_init(&dempty[1], 4 * NW);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  if (cl > 1) cluster_sync_all();  // the peers' barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t a_col0 = (FL ? NBUF : 1) * p.dcols;  // first TMEM column of the A-operand ring

  if (warp == 0) {
    // ===================================================================== TMA producer
    // the whole warp walks the ring; one elected lane issues
    const uint32_t tx = XB + bbytes * BPART * KB + ((SIDE_T && !H16) ? KB * 256 : 0);
    const int tile0i = (int)tile0;
    Pipe pp;
    for (int c = 0; c < nchunks; ++c, pp.advance(stages)) {
      const int st = pp.st;
      mbar_wait(&empty[st], pp.ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full[st], tx);
        uint8_t* b = bs + (size_t)st * BPART * KB * bbytes;
        const int brow = (chunk0 + c) * KB * (lp >> 3);
        if (!SIDE_T) {
          tma_load_2d(xs + (size_t)st * XB, &mapX, tile0i, c * KEL, &full[st], HINT_EVICT_FIRST);
        } else {
          const int k0 = (chunk0 + c) * KEL * KB;
          tma_load_2d(xs + (size_t)st * XB, &mapX, k0, tile0i, &full[st], HINT_EVICT_FIRST);
          if (!H16) bulk_load_1d(pd + st * KB * 256, p.pivot + 2 * k0, KB * 256, &full[st]);
        }
        // the images of consecutive slabs are contiguous in global memory: one bulk copy brings the stage's operand
        // (packed: [hi rows | lo rows] per slab, one copy for both parts) instead of lp/8 tensor-map rows of 1 KB —
        // the TMA engine's cost is per piece
        const size_t slab0 = (size_t)(chunk0 + c) * KB;
        if (cl > 1) {
          // this CTA's share of the image, written into every CTA of the cluster (same offsets, their barriers)
          const uint32_t nb0 = (PK ? 2 : 1) * KB * bbytes, share = nb0 / cl, off = share * cluster_ctarank();
          bulk_load_1d_mc(b + off, (const uint8_t*)(p.bimg_hi + slab0 * (PK ? 2 : 1) * lp * TC_KC) + off, share, &full[st],
                          cl_mask, HINT_EVICT_LAST);
          if (NS == 3 && !PK) {
            const uint32_t sh2 = KB * bbytes / cl, of2 = sh2 * cluster_ctarank();
            bulk_load_1d_mc(b + (size_t)KB * bbytes + of2, (const uint8_t*)(p.bimg_lo + slab0 * lp * TC_KC) + of2, sh2,
                            &full[st], cl_mask, HINT_EVICT_LAST);
          }
        } else if (p.b_bulk) {
          bulk_load_1d_hint(b, p.bimg_hi + slab0 * (PK ? 2 : 1) * lp * TC_KC, (PK ? 2 : 1) * KB * bbytes, &full[st],
                            HINT_EVICT_LAST);
          if (NS == 3 && !PK)
            bulk_load_1d_hint(b + (size_t)KB * bbytes, p.bimg_lo + slab0 * lp * TC_KC, KB * bbytes, &full[st], HINT_EVICT_LAST);
        } else {
          tma_load_2d(b, &mapBhi, 0, PK ? 2 * brow : brow, &full[st], HINT_EVICT_LAST);
          if (NS == 3 && !PK) tma_load_2d(b + (size_t)KB * bbytes, &mapBlo, 0, brow, &full[st], HINT_EVICT_LAST);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    {
      const uint32_t idesc = H16 ? make_idesc_f16(lp) : make_idesc(lp);
      const uint32_t idesc2 = make_idesc(2 * lp);  // packed: N = 2 lp
      Pipe pp;
      int g = 0, cg = 0;  // flush group and stage within it (NS == 3)
      for (int c = 0; c < nchunks; ++c, pp.advance(stages)) {
        const int st = pp.st;
        uint32_t d_tmem = tmem_base;
        bool first = c == 0;
        if (FL) {
          const int buf = g % NBUF;
          first = cg == 0;
          if (first) mbar_wait(&dempty[buf], (((uint32_t)g / NBUF) & 1) ^ 1);  // registers hold what this buffer had
          d_tmem = tmem_base + buf * p.dcols;
        }
        mbar_wait(&full[st], pp.ph);
        mbar_wait(&aready[st], pp.ph);
        tc_fence_after();
        const bool group_end = FL && (cg + 1 == FLUSH_STAGES || c == nchunks - 1);
        if (elect_one()) {
          const uint32_t a_hi = tmem_base + a_col0 + st * ACOLS;
          const uint32_t b0 = smem_u32(bs + (size_t)st * BPART * KB * bbytes);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const uint64_t dh = make_b_desc(b0 + kb * (PK ? 2 : 1) * bbytes);
            const uint64_t dl = (NS == 3 && !PK) ? make_b_desc(b0 + (KB + kb) * bbytes) : 0;
#pragma unroll
            for (int k = 0; k < TC_KC / 8; ++k) {
              // +32 bytes (8 tf32) along K inside the swizzle atom = +2 in the (address >> 4) field
              const uint32_t a = a_hi + kb * TC_KC + k * 8;
              if (PK) {
                mma_tf32_ts(d_tmem, a, dh + 2 * k, idesc2, !(first && kb == 0 && k == 0));  // a_hi [b_hi | b_lo]
                mma_tf32_ts(d_tmem, a + TC_KC * KB, dh + 2 * k, idesc, 1);                  // a_lo b_hi
              } else if (H16) {
                mma_f16_ts(d_tmem, a, dh + 2 * k, idesc, !(first && kb == 0 && k == 0));  // K = 16: 8 columns of pairs
              } else {
                mma_tf32_ts(d_tmem, a, dh + 2 * k, idesc, !(first && kb == 0 && k == 0));
                if (NS == 3) mma_tf32_ts(d_tmem, a, dl + 2 * k, idesc, 1);
                if (NS >= 2) mma_tf32_ts(d_tmem, a + TC_KC * KB, dh + 2 * k, idesc, 1);
              }
            }
          }
          if (cl > 1) mma_commit_mc(&empty[st], cl_mask);
          else mma_commit(&empty[st]);
          if (group_end) mma_commit(&dfull[g % NBUF]);
          if (!FL && c == nchunks - 1) mma_commit(&dfull[0]);
        }
        __syncwarp();
        if (FL) {
          if (group_end) { ++g; cg = 0; } else { ++cg; }
        }
      }
    }
  } else {
    // ===================================================================== operand stage + epilogue
    // NW warps share each TMEM lane quarter: warp `part` of a quarter converts KW = 32/NW of the 32 K values of a
    // slab and owns lp/NW columns of the accumulator
    const int q = warp & 3;             // TMEM lane quarter this warp may touch
    const int part = (warp - 2) >> 2;
    const int row = q * 32 + lane;      // row of D / TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float piv = 0.f;
    if (!SIDE_T && !STATS && !H16) piv = (tile0 + row < p.S) ? p.pivot[tile0 + row] : 0.f;
    // fused statistics: the first sample is the shift of the sums and the pivot of this pass
    bool n0 = false;           // the first sample of this feature is NaN
    double sum1 = 0.0, sum2 = 0.0;
    int cnt = 0;
    float c0 = 0.f;            // statistics pass that also writes the fp16 copy: this feature's power of two
    if (STATS && MODE == MODE_WCOPY) c0 = (tile0 + row < p.S) ? p.c0[tile0 + row] : 0.f;
    if (STATS) {
      const float x0 = (tile0 + row < p.S) ? p.X[tile0 + row] : 0.f;
      n0 = !(x0 == x0);
      piv = n0 ? 0.f : x0;
      if (part == 0) {
        const unsigned b0 = __ballot_sync(0xffffffffu, n0);
        if (lane == 0 && b0) atomicAdd(p.base_nan, __popc(b0));
      }
    }

    constexpr int ACCN = FL ? 128 / NW : 1;
    const int cw = lp / NW;      // accumulator columns of this warp: [part*cw, +cw)
    const int groups = cw >> 2;  // in groups of 4
    float acc[ACCN];
#pragma unroll
    for (int i = 0; i < ACCN; ++i) acc[i] = 0.f;
    const int nflush = (nchunks + FLUSH_STAGES - 1) / FLUSH_STAGES;
    int next_flush = 0;
    // registers += accumulator buffer of flush group g (after the MMA warp committed it), then hand the buffer back
    auto flush = [&](int g) {
      const int buf = g % NBUF;
      mbar_wait(&dfull[buf], ((uint32_t)g / NBUF) & 1);
      tc_fence_after();
#pragma unroll
      for (int gi = 0; gi < ACCN / 4; ++gi) {
        if (gi < groups) {
          float v[4];
          tmem_ld4(tmem_base + lane_addr + buf * p.dcols + part * cw + gi * 4, v);
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[(FL ? gi * 4 + e : 0)] += v[e];
          if (PK) {  // the a_hi b_lo half
            tmem_ld4(tmem_base + lane_addr + lp + part * cw + gi * 4, v);
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[(FL ? gi * 4 + e : 0)] += v[e];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dempty[buf]);
    };

    const uint32_t xs_u32 = smem_u32(xs), pd_u32 = smem_u32(pd);
    Pipe pp;
    for (int c = 0; c < nchunks; ++c, pp.advance(stages)) {
      const int st = pp.st;
      if (FL && next_flush < nflush && c >= (next_flush + 1) * FLUSH_STAGES + 1) flush(next_flush++);
      // project_S: a NaN only spoils the row of D of its own feature (dropped in the epilogue if the feature is
      // invalid), except in samples that are NaN throughout: only stages holding such a sample test every value
      const bool check = SIDE_T || (p.chunk_flags != nullptr && p.chunk_flags[c] != 0);
      mbar_wait(&full[st], pp.ph);
      tc_fence_after();
      const uint32_t a_slot = tmem_base + lane_addr + a_col0 + st * ACOLS + part * KW;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        uint32_t hi[KW], lo[KW];
        if (!SIDE_T && H16) {
          // X stage = [64 t][128 s] fp16; this thread owns column `row`: TMEM column i of the slab holds the pair
          // (t = 2i, 2i + 1), this warp converts columns part*KW .. +KW-1
          const uint32_t src = xs_u32 + st * XB + (part * KW * 2) * (TC_TILE * 2) + row * 2;
#pragma unroll
          for (int r = 0; r < KW; ++r)
            hi[r] = lds16(src + (2 * r) * (TC_TILE * 2)) | (lds16(src + (2 * r + 1) * (TC_TILE * 2)) << 16);
        } else if (!SIDE_T) {
          // X stage = [32 t][128 s] fp32; this thread owns column `row`, rows part*KW .. +KW-1
          const uint32_t src = xs_u32 + st * XB + ((part * KW) * TC_TILE + row) * 4;
          float v[KW];
#pragma unroll
          for (int r = 0; r < KW; ++r) v[r] = lds32(src + r * TC_TILE * 4) - piv;
          if (STATS) {
            // sums of (x - shift), (x - shift)^2 and the count over the samples of this stage (fp32 within the
            // stage, fp64 across stages).  A feature without NaN in the stage — seen from the sum itself — takes the
            // short road; the per-value tests run only for NaN features, the ragged last stage and NaN first samples.
            const int64_t t0 = (int64_t)c * TC_KC + part * KW;
            const bool ragged = (c == nchunks - 1) && (p.T & (TC_KC - 1));
            float s1 = 0.f, s2 = 0.f;
            int cn = KW;
            uint32_t devmask = 0;  // samples whose NaN flag differs from the first sample's
#pragma unroll
            for (int r = 0; r < KW; ++r) s1 += v[r];
            if (!ragged && !n0 && fabsf(s1) <= 3.4028234e38f) {
#pragma unroll
              for (int r = 0; r < KW; ++r) s2 = fmaf(v[r], v[r], s2);
            } else {
              s1 = 0.f;
              cn = 0;
#pragma unroll
              for (int r = 0; r < KW; ++r) {
                const bool live = !ragged || t0 + r < p.T;
                const bool ok = v[r] == v[r];
                const float dz = (ok && live) ? v[r] : 0.f;
                s1 += dz;
                s2 = fmaf(dz, dz, s2);
                cn += (ok && live) ? 1 : 0;
                v[r] = dz;
                if (live && ((!ok) != n0) && (tile0 + row < p.S)) devmask |= 1u << r;
              }
            }
            if (__any_sync(0xffffffffu, devmask != 0)) {
#pragma unroll
              for (int r = 0; r < KW; ++r) {
                const unsigned dev = __ballot_sync(0xffffffffu, (devmask >> r) & 1);
                if (dev) {
                  const unsigned pos = __ballot_sync(0xffffffffu, ((devmask >> r) & 1) && !n0);
                  if (lane == 0) atomicAdd(&p.row_delta[t0 + r], 2 * __popc(pos) - __popc(dev));
                }
              }
            }
            sum1 += (double)s1;
            sum2 += (double)s2;
            cnt += cn;
            if (MODE == MODE_WCOPY && tile0 + row < p.S) {
              // A16[t, s] = fp16((x - first sample) c0): 32 lanes = 64 contiguous bytes of a row of the copy
              uint16_t* dst = p.copy16 + t0 * p.ldc16 + tile0 + row;
#pragma unroll
              for (int r = 0; r < KW; ++r)
                if (!ragged || t0 + r < p.T) dst[(int64_t)r * p.ldc16] = (uint16_t)(pack_h2(v[r] * c0, 0.f) & 0xffffu);
            }
          } else if (check) {
#pragma unroll
            for (int r = 0; r < KW; ++r) v[r] = (v[r] == v[r]) ? v[r] : 0.f;
          }
#pragma unroll
          for (int r = 0; r < KW; ++r) {
            if (NS >= 2) {
              hi[r] = __float_as_uint(v[r]) & 0xffffe000u;
              lo[r] = __float_as_uint(v[r] - __uint_as_float(hi[r]));  // the hardware truncates it: -2^-22, see DESIGN.md
            } else {
              hi[r] = RN == 1 ? to_tf32(v[r]) : __float_as_uint(v[r]);  // the tensor core reads the upper 19 bits
            }
          }
        } else {
          // X stage = 128 rows (t) of KB*32 (+4 unused) fp32, XPITCH bytes apart; this thread owns row `row`, 16-byte
          // chunks part*KW/4 .. of slab kb.  Anything that is not a finite number counts as 0 (NaN samples /
          // features; beyond the last feature dscale is 0 and the small operand too)
          const uint32_t src = xs_u32 + st * XB + row * XPITCH + kb * 128;
          const uint32_t pv = pd_u32 + st * KB * 256 + kb * 256;
          constexpr bool FOLD = NS != 2;  // dscale lives in the small operand's image
          uint32_t cw[KW / 2 > 0 ? KW / 2 : 1];  // MODE_WCOPY: the packed fp16 pairs of this thread's values
#pragma unroll
          for (int cc = 0; cc < KW / 4; ++cc) {
            const int ch = part * (KW / 4) + cc;
            const float4 x = lds128(src + ch * 16);
            if (RN == 2 || H16) {  // materialised field (rounded fp32 copy, or packed fp16 pairs): a plain move
              hi[cc * 4 + 0] = __float_as_uint(x.x); hi[cc * 4 + 1] = __float_as_uint(x.y);
              hi[cc * 4 + 2] = __float_as_uint(x.z); hi[cc * 4 + 3] = __float_as_uint(x.w);
              continue;
            }
            const float4 pq = lds128(pv + ch * 16);
            float4 dq = make_float4(1.f, 1.f, 1.f, 1.f);
            if (!FOLD || MODE == MODE_WCOPY) dq = lds128(pv + 128 + ch * 16);  // (WCOPY: the slot carries e16)
            const float xa[4] = {x.x, x.y, x.z, x.w}, pa[4] = {pq.x, pq.y, pq.z, pq.w}, da[4] = {dq.x, dq.y, dq.z, dq.w};
            float wv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float v = xa[e] - pa[e];
              if (!FOLD) v *= da[e];
              if (!TFAST) v = (fabsf(v) <= 3.4028234e38f) ? v : 0.f;
              if (MODE == MODE_WCOPY) wv[e] = v;
              if (NS >= 2) {
                hi[cc * 4 + e] = __float_as_uint(v) & 0xffffe000u;
                lo[cc * 4 + e] = __float_as_uint(v - __uint_as_float(hi[cc * 4 + e]));
              } else {
                hi[cc * 4 + e] = RN == 1 ? to_tf32(v) : __float_as_uint(v);
              }
            }
            if (MODE == MODE_WCOPY) {
              cw[cc * 2] = pack_h2(wv[0] * da[0], wv[1] * da[1]);
              cw[cc * 2 + 1] = pack_h2(wv[2] * da[2], wv[3] * da[3]);
            }
          }
          if (MODE == MODE_WCOPY) {
            // this thread's KW values of the slab: KW * 2 contiguous bytes of its row of the staging tile
            const uint32_t dst = smem_u32(cst) + (c & 1) * (TC_TILE * CPITCH) + row * CPITCH + kb * 64 + part * KW * 2;
#pragma unroll
            for (int i = 0; i < KW / 8; ++i)
              sts128(dst + i * 16, make_float4(__uint_as_float(cw[i * 4]), __uint_as_float(cw[i * 4 + 1]),
                                               __uint_as_float(cw[i * 4 + 2]), __uint_as_float(cw[i * 4 + 3])));
          }
        }
        tmem_st<KW>(a_slot + kb * TC_KC, hi);
        if (NS >= 2) tmem_st<KW>(a_slot + kb * TC_KC + TC_KC * KB, lo);
      }
      if (MODE == MODE_WCOPY && SIDE_T) {
        // all operand warps have written the stage's tile: every warp stores 16 of its rows, 4 rows (4 x KB*64 bytes
        // of contiguous global memory each) per instruction
        asm volatile("bar.sync 1, %0;" ::"r"(128 * NW) : "memory");
        const uint32_t tile = smem_u32(cst) + (c & 1) * (TC_TILE * CPITCH);
        constexpr int CHUNKS = KB * 4;                 // 16-byte chunks per row
        constexpr int RPI = 32 / CHUNKS;               // rows per instruction
        const int w8 = warp - 2;                       // 0 .. 4*NW-1
        const int rows_per_warp = TC_TILE / (4 * NW);
        const int64_t s0 = ((int64_t)(chunk0 + c) * KB) * TC_KC;
#pragma unroll
        for (int i = 0; i < rows_per_warp / RPI; ++i) {
          const int r = w8 * rows_per_warp + i * RPI + lane / CHUNKS, ch = lane % CHUNKS;
          const float4 v = lds128(tile + r * CPITCH + ch * 16);
          if (tile0 + r < p.T)
            *reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(p.copy16 + (tile0 + r) * p.ldc16 + s0) + ch * 16) = v;
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&aready[st]);
    }

    // ---- epilogue: D (128 x lp fp32: TMEM for x1, registers for x3) -> global
    if (FL) {
      while (next_flush < nflush) flush(next_flush++);
    } else {
      mbar_wait(&dfull[0], 0);
      tc_fence_after();
    }
    const int64_t rrow = tile0 + row;  // s (project_S) or t (project_T)
    const bool ok = SIDE_T ? true : rrow < p.S;
    float ds = 0.f, cs = 0.f;
    if (STATS) {
      // the NW warps of a lane quarter each saw a share of the samples: combine, then every thread derives the
      // Scaler vectors of its feature (scaling_finalize_kernel's arithmetic, stats.cu)
      __shared__ double comb1[2][TC_TILE], comb2[2][TC_TILE];
      __shared__ int combn[2][TC_TILE];
      comb1[part][row] = sum1;
      comb2[part][row] = sum2;
      combn[part][row] = cnt;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const double t1 = comb1[0][row] + comb1[1][row], t2 = comb2[0][row] + comb2[1][row];
      const int n = combn[0][row] + combn[1][row];
      const bool center = p.stat_flags & XEOFS_F_CENTER, standardize = p.stat_flags & XEOFS_F_STANDARDIZE;
      const bool val = ok && n > 0;
      double mu = 0.0, m2 = 0.0;
      if (val) {
        const double a = t1 / n;
        mu = (double)piv + a;
        m2 = t2 - t1 * a;
        if (m2 < 0) m2 = 0;
      }
      const float mu32 = (float)mu;
      const float sd = val ? fmaxf((float)sqrt(m2 / (n > 0 ? n : 1)), 1.1920929e-07f) : nanf("");
      double d = val ? (p.featw ? p.featw[rrow] : 1.0) : 0.0;
      if (standardize && val) d /= (double)sd;
      const float d32 = (float)d;
      const float mu_eff = center ? mu32 : 0.f;
      ds = d32;
      cs = val ? (piv - mu_eff) * d32 : 0.f;   // A = (x - shift) d + (shift - mean_eff) d
      double tv = (val && n > 1) ? (double)d32 * (double)d32 * m2 / (double)(n - 1) : 0.0;
      int nv = val ? 1 : 0, cmax = val ? n : 0, cmin = val ? n : 0x7fffffff;
      if (part == 0 && ok) {
        const float newpiv = val ? (center ? mu32 : piv) : 0.f;
        p.mean_out[rrow] = val ? mu32 : nanf("");
        p.std_out[rrow] = sd;
        p.valid_out[rrow] = val ? 1 : 0;
        p.pivot_out[rrow] = newpiv;
        p.dscale_out[rrow] = d32;
        p.ccorr_out[rrow] = val ? (newpiv - mu_eff) * d32 : 0.f;
        if (MODE == MODE_WCOPY) {
          p.ic16_out[rrow] = (val && c0 != 0.f) ? d32 / c0 : 0.f;
          p.cc16_out[rrow] = cs;  // (first sample - mean_eff) dscale: what the shifted copy lacks
        }
      }
      if (part != 0) { tv = 0.0; nv = 0; cmax = 0; cmin = 0x7fffffff; }
      tv = warp_sum(tv);
      nv = warp_sum(nv);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
        cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
      }
      if (part == 0 && lane == 0) {
        atomicAdd(&p.scalars_out[0], tv);
        atomicAdd(&p.scalars_out[1], (double)nv);
        atomicMax((unsigned long long*)&p.scalars_out[2], (unsigned long long)__double_as_longlong((double)cmax));
        atomicMin((unsigned long long*)&p.scalars_out[3], (unsigned long long)__double_as_longlong((double)cmin));
      }
    } else if (!SIDE_T && ok) {
      ds = p.dscale[rrow];
      cs = p.ccorr ? p.ccorr[rrow] : 0.f;
    }
    float* dstT = SIDE_T ? p.out + ((int64_t)blockIdx.y * gridDim.x * TC_TILE + rrow) * lp : nullptr;
#pragma unroll
    for (int gi = 0; gi < 128 / NW / 4; ++gi) {
      if (gi < groups) {
        const int j0 = part * cw + gi * 4;
        float v[4];
        if (FL) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = acc[(FL ? gi * 4 + e : 0)];
        } else {
          tmem_ld4(tmem_base + lane_addr + j0, v);
        }
        if (!SIDE_T) {
          if (ok) {
            // an invalid feature (dscale 0) gives a zero row even if its accumulator holds NaN
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (H16) v[e] *= p.igs[j0 + e];  // undo the power-of-two column scale of the fp16 operand image
              p.out[(int64_t)(j0 + e) * p.ldo + rrow] =
                  ds != 0.f ? fmaf(ds, v[e], (STATS || p.ccorr) ? cs * p.wsum[j0 + e] : 0.f) : 0.f;
            }
          }
        } else {
          *reinterpret_cast<float4*>(dstT + j0) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
  if (cl > 1) cluster_sync_all();  // no CTA leaves while a peer may still write into its shared memory / barriers
}

// ------------------------------------------------------------------------------------------------ small helpers

// W (T x ldw, time-side) -> images of the K slabs of W^T (K = t): hi = value with the TF32 bits only, lo = fp32
// remainder (3xTF32 only, else the value goes in unsplit).  Zero beyond T.
__global__ void __launch_bounds__(256)
prep_W_kernel(const float* __restrict__ W, int64_t T, int64_t ldw, int lp, float* __restrict__ Whi, float* __restrict__ Wlo,
              int rn = 0, int slab_floats = 0) {
  // one block per slab of 32 t: thread (j, 4 k's).  slab_floats: distance between the images of consecutive slabs
  // (lp * 32, or 2 * lp * 32 when hi and lo rows share one image: Wlo = Whi + lp * 32)
  const int64_t t0 = (int64_t)blockIdx.x * 32;
  const size_t ss = slab_floats ? (size_t)slab_floats : (size_t)lp * 32;
  float* hi = Whi + (size_t)blockIdx.x * ss;
  float* lo = Wlo ? Wlo + (size_t)blockIdx.x * ss : nullptr;
  for (int idx = threadIdx.x; idx < lp * 32; idx += 256) {
    const int kk = idx / lp, j = idx % lp;  // consecutive threads walk a row of W
    const int64_t t = t0 + kk;
    const float v = t < T ? W[t * ldw + j] : 0.f;
    if (lo) {
      const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
      hi[img_offset(j, kk)] = h;
      lo[img_offset(j, kk)] = __uint_as_float(to_tf32(v - h));
    } else {
      hi[img_offset(j, kk)] = rn ? __uint_as_float(to_tf32(v)) : v;
    }
  }
}

// pivot / dscale zero-padded to Spad and interleaved per 32-wide slab ([slab][pivot 32 | dscale 32]) so that one bulk
// copy brings both for a stage of project_T
__global__ void pad_vectors_kernel(const float* __restrict__ pivot, const float* __restrict__ dscale, int64_t S,
                                   int64_t Spad, float* __restrict__ pdpad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Spad) {
    const int64_t o = (i >> 5) * 64 + (i & 31);
    pdpad[o] = i < S ? pivot[i] : 0.f;
    pdpad[o + 32] = i < S ? dscale[i] : 0.f;
  }
}

// flags[c] = 1 if one of the 32 samples of K-chunk c is NaN throughout (row_valid == 0)
__global__ void chunk_flags_kernel(const uint8_t* __restrict__ row_valid, int64_t T, int nchunks, uint8_t* __restrict__ flags) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  uint8_t f = 0;
  for (int r = 0; r < 32; ++r) {
    const int64_t t = (int64_t)c * 32 + r;
    if (t < T && !row_valid[t]) f = 1;
  }
  flags[c] = f;
}

// Yt (lp x ldy, space-side) -> images of its K slabs (K = s), zero beyond S.  One block per slab.
__global__ void __launch_bounds__(256)
tile_Y_kernel(const float* __restrict__ Yt, int64_t S, int64_t ldy, int lp, float* __restrict__ Yhi, float* __restrict__ Ylo,
              int rn = 0, int slab_floats = 0, const float* __restrict__ dscale = nullptr) {
  const int64_t s0 = (int64_t)blockIdx.x * 32;
  const size_t ss = slab_floats ? (size_t)slab_floats : (size_t)lp * 32;
  float* hi = Yhi + (size_t)blockIdx.x * ss;
  float* lo = Ylo ? Ylo + (size_t)blockIdx.x * ss : nullptr;
  const int kk = threadIdx.x & 31;
  for (int j = threadIdx.x >> 5; j < lp; j += 8) {
    // dscale (the Scaler's per-feature factor, 0 for dropped features) rides on the small operand where given
    const float ds = (dscale && s0 + kk < S) ? dscale[s0 + kk] : 1.f;
    const float v = (s0 + kk < S) ? Yt[(int64_t)j * ldy + s0 + kk] * ds : 0.f;
    if (lo) {
      const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
      hi[img_offset(j, kk)] = h;
      lo[img_offset(j, kk)] = __uint_as_float(to_tf32(v - h));
    } else {
      hi[img_offset(j, kk)] = rn ? __uint_as_float(to_tf32(v)) : v;
    }
  }
}

// Z[t, j] = sum_split P[split][t][j] + r[j]   (r only on the valid samples)
__global__ void reduce_partials_kernel(const float* __restrict__ P, int splits, int64_t rows_pad, int lp, int64_t T,
                                       const float* __restrict__ r, const uint8_t* __restrict__ row_valid,
                                       float* __restrict__ Z, int64_t ldz, const float* __restrict__ colfac = nullptr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * lp) return;
  const int64_t t = i / lp;
  const int j = (int)(i % lp);
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += P[((int64_t)s * rows_pad + t) * lp + j];
  if (colfac) acc *= colfac[j];
  if (r && (!row_valid || row_valid[t])) acc += r[j];
  Z[t * ldz + j] = acc;
}

// ------------------------------------------------------------------------------------------------ fp16 path helpers
// Offset (in halves) of element (row j, k) inside the fp16 image of one 64-wide K slab: [rows][64 k], 16-byte chunks
// (8 halves) XOR-swizzled by (row & 7) — the K-major SWIZZLE_128B layout.
__device__ __forceinline__ int img16_offset(int j, int kk) { return j * 64 + ((((kk >> 3) ^ (j & 7)) << 3) | (kk & 7)); }

// e16[s] = dscale[s] c[s], ic16[s] = 1 / c[s] with c = 2^round(log2(512 / (|dscale| std))): the copy's columns get a
// standard deviation of ~512 (fp16: +-65504 = 127 sigma, full 11 bits down to 1e-7 sigma); dropped features: 0
__global__ void h16_scales_kernel(const float* __restrict__ dscale, const float* __restrict__ stdv, int64_t S,
                                  float* __restrict__ e16, float* __restrict__ ic16) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S) return;
  const float d = dscale[i], sd = stdv[i];
  float e = 0.f, ic = 0.f;
  if (d != 0.f && sd == sd && sd > 0.f) {
    int k = (int)rintf(log2f(512.f / (fabsf(d) * sd)));
    k = max(-60, min(60, k));
    e = ldexpf(d, k);
    ic = ldexpf(1.f, -k);
  }
  e16[i] = e;
  ic16[i] = ic;
}

// c0[s] = the power of two that puts the largest |x - first sample| met in 64 samples spread over the record at 512
// (fp16 then holds deviations 128 times larger before it saturates)
__global__ void h16_prescale_kernel(const float* __restrict__ X, int64_t T, int64_t S, int64_t ldx, float* __restrict__ c0) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float x0 = X[s];
  if (!(x0 == x0)) x0 = 0.f;
  float amax = 0.f;
  for (int i = 1; i <= 64; ++i) {
    const int64_t t = (i * T) / 65;
    const float d = X[t * ldx + s] - x0;
    if (d == d) amax = fmaxf(amax, fabsf(d));
  }
  if (!(amax > 0.f) || !(amax < 3.0e38f)) amax = fmaxf(fabsf(x0) * 2.44e-4f, 1e-30f);
  int k = (int)floorf(log2f(512.f / amax));
  k = max(-100, min(100, k));
  c0[s] = ldexpf(1.f, k);
}

// amax[j] = max_n |M(n, j) f(n)| over a k-column matrix (side 0: time-side n x ld, f = 1; side 1: space-side, f = fac).
// Side 1 with cc also leaves the block's share of r[j] = sum_n cc[n] M(n, j) in dpart[blockIdx.x * lp + j]
// (h16_colscale_kernel adds the shares in block order, so r does not depend on scheduling).
__global__ void __launch_bounds__(256)
absmax_cols_kernel(const float* __restrict__ M, int64_t n, int64_t ld, int lp, int side, const float* __restrict__ fac,
                   float* __restrict__ amax, const float* __restrict__ cc, double* __restrict__ dpart) {
  if (side == 1) {
    const int j = blockIdx.y;
    float m = 0.f;
    double acc = 0.0;
    const float* row = M + (int64_t)j * ld;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const float v = row[i];
      m = fmaxf(m, fabsf(v * (fac ? fac[i] : 1.f)));
      if (cc) acc += (double)cc[i] * (double)v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(&amax[j]), __float_as_int(m));
    if (cc) {
      __shared__ double sh[8];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
      __syncthreads();
      if (threadIdx.x == 0) {
        double a = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += sh[w];
        dpart[(int64_t)blockIdx.x * lp + j] = a;
      }
    }
  } else {
    // consecutive threads walk a row; thread t owns column t % lp for rows t / lp + i * (256 / lp)... simple version:
    for (int j = threadIdx.x; j < lp; j += blockDim.x) {
      float m = 0.f;
      for (int64_t i = blockIdx.x; i < n; i += gridDim.x) m = fmaxf(m, fabsf(M[i * ld + j]));
      if (m > 0.f) atomicMax(reinterpret_cast<int*>(&amax[j]), __float_as_int(m));
    }
  }
}
__global__ void recip_kernel(const float* __restrict__ a, int n, float* __restrict__ o) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) o[j] = 1.f / a[j];
}
// gs[j] = the power of two that brings amax[j] to ~2^12 (1 for an empty column); r[j] = the sum of the nparts shares
__global__ void h16_colscale_kernel(const float* __restrict__ amax, int lp, float* __restrict__ gs,
                                    const double* __restrict__ dpart, int nparts, float* __restrict__ r) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= lp) return;
  const float a = amax[j];
  int k = 0;
  if (a > 0.f && a == a && a < 3.0e38f) k = max(-100, min(100, 12 - (int)ceilf(log2f(a))));
  gs[j] = ldexpf(1.f, k);
  if (dpart) {
    double acc = 0.0;
    for (int b = 0; b < nparts; ++b) acc += dpart[(int64_t)b * lp + j];
    r[j] = (float)acc;
  }
}
// W (T x ldw, time-side) -> fp16 images of the 64-wide K slabs of W^T (K = t), column j scaled by gs[j].  Zero beyond T.
__global__ void __launch_bounds__(256)
prep_W16_kernel(const float* __restrict__ W, int64_t T, int64_t ldw, int lp, const float* __restrict__ gs,
                uint16_t* __restrict__ img) {
  const int64_t t0 = (int64_t)blockIdx.x * 64;
  uint16_t* o = img + (size_t)blockIdx.x * lp * 64;
  for (int idx = threadIdx.x; idx < lp * 64; idx += 256) {
    const int kk = idx / lp, j = idx % lp;  // consecutive threads walk a row of W
    const int64_t t = t0 + kk;
    const float v = t < T ? W[t * ldw + j] * gs[j] : 0.f;
    o[img16_offset(j, kk)] = (uint16_t)(pack_h2(v, 0.f) & 0xffffu);
  }
}
// Yt (lp x ldy, space-side) -> fp16 images of its 64-wide K slabs (K = s): Yt[j,s] ic16[s] gs[j].  Zero beyond S.
__global__ void __launch_bounds__(256)
tile_Y16_kernel(const float* __restrict__ Yt, int64_t S, int64_t ldy, int lp, const float* __restrict__ ic16,
                const float* __restrict__ gs, uint16_t* __restrict__ img) {
  const int64_t s0 = (int64_t)blockIdx.x * 64;
  uint16_t* o = img + (size_t)blockIdx.x * lp * 64;
  const int kk = threadIdx.x & 63;
  const float f = (s0 + kk < S) ? ic16[s0 + kk] : 0.f;
  for (int j = threadIdx.x >> 6; j < lp; j += 4) {
    const float v = (s0 + kk < S) ? Yt[(int64_t)j * ldy + s0 + kk] * f * gs[j] : 0.f;
    o[img16_offset(j, kk)] = (uint16_t)(pack_h2(v, 0.f) & 0xffffu);
  }
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

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

// fp32 tensor map of rank 2 (inner, outer) or rank 3 (inner, mid, outer); strides in elements
// swizzle: 0 none, 1 = 128-byte (16-byte chunks), 2 = 128-byte with 32-byte atoms (the MN-major image of 32-bit operands)
static int make_map(CUtensorMap* m, const float* base, int rank, const int64_t* dims, const int64_t* strides,
                    const int* box, int swizzle) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return XEOFS_E_UNSUPPORTED;
  }
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t bx[3], estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) { gdim[i] = (cuuint64_t)dims[i]; bx[i] = (cuuint32_t)box[i]; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = (cuuint64_t)strides[i] * sizeof(float);
  static const CUtensorMapL2promotion promo[4] = {CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_64B,
                                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base, gdim, gstr, bx, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle == 2   ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                   : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : CU_TENSOR_MAP_SWIZZLE_NONE,
                   promo[env_int("XEOFS_TC_PROMO", 3) & 3], CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d extent %lld x %lld box %d x %d", (int)r, rank, (long long)dims[0],
              (long long)dims[1], box[0], box[1]);
    return XEOFS_E_CUDA;
  }
  return XEOFS_OK;
}
int make_map2(CUtensorMap* m, const float* base, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                     int box_outer, int swizzle) {
  const int64_t dims[2] = {inner, outer}, str[1] = {ld};
  const int box[2] = {box_inner, box_outer};
  return make_map(m, base, 2, dims, str, box, swizzle);
}

// 16-bit tensor map of rank 2 (the fp16 copy of the preprocessed matrix), no swizzle, OOB -> 0
static int make_map16(CUtensorMap* m, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return XEOFS_E_UNSUPPORTED;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer}, gstr[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t bx[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer}, estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, (void*)base, gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (16-bit) failed (%d): extent %lld x %lld box %d x %d", (int)r, (long long)inner,
              (long long)outer, box_inner, box_outer);
    return XEOFS_E_CUDA;
  }
  return XEOFS_OK;
}

bool tensor_maps_available() { return get_encode() != nullptr; }

bool tc_supported(int64_t T, int64_t S, int64_t ldx, const float* X, int64_t l) {
  (void)T; (void)S;
  return l >= 1 && l <= 128 && ldx % 4 == 0 && ((uintptr_t)X % 16 == 0) && get_encode() != nullptr;
}

static inline int64_t align256(int64_t b) { return round_up(b, 256); }
static inline bool is_x3(int algo) { return algo == XEOFS_ALGO_TF32X3 || algo == XEOFS_ALGO_AUTO; }
static inline int algo_ns(int algo) { return algo == XEOFS_ALGO_TF32X3 ? 3 : algo == XEOFS_ALGO_TF32X2 ? 2 : 1; }
static inline int algo_rn(int algo) { return algo == XEOFS_ALGO_TF32X1R ? 1 : algo == XEOFS_ALGO_TF32X1F ? 2 : 0; }
// 3xTF32 with the two products of the big operand's upper part packed into one N = 2 lp instruction (see the kernel)
static inline bool use_pack(int ns, int lp) { return ns == 3 && lp <= 96 && env_int("XEOFS_TC_PACK", 1) != 0; }

// project_T: 32-wide K slabs per stage = bytes of a row fetched per bulk copy / 128.  The largest that leaves
// at least two stages of shared memory and TMEM.
struct Shape {
  int stages, dcols, kb;
  uint32_t tmem_cols;
  size_t smem;
};
static Shape pick_shape(int lp, int ns, bool side_t, int kb, bool rn = false, bool wcopy = false) {
  Shape sh;
  const int npart = ns >= 2 ? 2 : 1, bpart = ns == 3 ? 2 : 1;
  const bool two_ctas = !side_t && ns == 1 && !rn;  // project_S x1: two CTAs per SM share the 512 TMEM columns
  const bool pk = use_pack(ns, lp);
  const int xb = side_t ? TC_TILE * (kb * 128 + 16) : TC_XBYTES;
  const int per_stage = xb + lp * TC_KC * 4 * bpart * kb + (side_t ? kb * 256 : 0);
  sh.kb = kb;
  sh.dcols = (int)round_up(pk ? 2 * lp : lp, 32);
  sh.tmem_cols = two_ctas ? 256 : 512;
  const int ring = (int)sh.tmem_cols - (pk ? 1 : (ns >= 2 || rn) ? 2 : 1) * sh.dcols;
  const int by_tmem = ring / (TC_KC * kb * npart);
  const int extra = wcopy ? 2 * TC_TILE * (kb * 64 + 16) : 0;  // the staging tiles of the pass that writes the fp16 copy
  const int budget = (two_ctas ? 110 : 222) * 1024 - 2048 - extra;
  int st = budget / per_stage;
  if (st > by_tmem) st = by_tmem;
  if (st > TC_MAX_STAGES) st = TC_MAX_STAGES;
  sh.stages = st;
  sh.smem = (size_t)(st > 0 ? st : 1) * per_stage + extra + 1024 /*alignment*/ + 256 /*barriers*/;
  return sh;
}
static Shape pick_shape_T(int lp, int ns, int64_t S, int64_t ldx, bool rn = false, bool wcopy = false) {
  int kb = env_int("XEOFS_TC_KB", 2);
  if (kb != 1 && kb != 2 && kb != 4) kb = 2;
  (void)S; (void)ldx;
  Shape sh = pick_shape(lp, ns, true, kb, rn, wcopy);
  while (sh.stages < 2 && kb > 1) {
    kb >>= 1;
    sh = pick_shape(lp, ns, true, kb, rn, wcopy);
  }
  return sh;
}

struct TGeom {
  int64_t t_tiles, rows_pad, Spad;
  int chunks_total, chunks_per_cta, splits;
};
// cap: the longest K range (in 32-wide slabs) one TMEM accumulator may sum before its truncating adds show: 0 = no cap
// (the 3xTF32 / 2xTF32 / rounded-TF32 kernels flush into fp32 registers every TC_FLUSH slabs), 1024 for the
// power-iteration products
// (a bias of ~1e-4 that only rescales the iterate)
static int t_cap(int algo) { return (algo_ns(algo) >= 2 || algo_rn(algo) != 0) ? 0 : 1024; }
// CTAs per cluster sharing the small operand's image through TMA multicast (XEOFS_TC_CLUSTER = 1, 2 or 4).
// Measured on B200 (profiles/r02_cluster_multicast_sweep.txt): clusters of 2 or 4 change nothing for project_S and cost
// project_T 5-10 % at lp = 64 .. 112 — the operand image is an L2 hit either way and ncu shows the L2 at 57 % of its
// peak in these passes, so the traffic multicast removes was never the limiter; the lock-step it adds between the
// CTAs of a cluster is what shows.  Default 1.
static int pick_cluster() {
  const int c = env_int("XEOFS_TC_CLUSTER", 1);
  return c == 4 ? 4 : c == 2 ? 2 : 1;
}
static TGeom t_geometry(int64_t T, int64_t S, int cap, int kb, int cl = 1, int kel = TC_KC) {
  TGeom g;
  g.t_tiles = round_up(ceil_div(T, TC_TILE), cl);  // whole clusters along the row tiles (the extra tiles hold no rows)
  g.rows_pad = g.t_tiles * TC_TILE;
  g.Spad = round_up(S, 128);  // covers every KB
  g.chunks_total = (int)ceil_div(S, kel * kb);  // kel: K values per slab (32 fp32, 64 fp16)
  int64_t want = ceil_div(2 * (int64_t)num_sms(), g.t_tiles);
  if (want < 1) want = 1;
  if (want > g.chunks_total) want = g.chunks_total;
  g.chunks_per_cta = (int)ceil_div(g.chunks_total, want);
  if (cap > 0 && g.chunks_per_cta > cap / kb) g.chunks_per_cta = cap / kb;
  g.splits = (int)ceil_div(g.chunks_total, g.chunks_per_cta);
  return g;
}

int64_t tc_workspace_bytes(int64_t T, int64_t S, int64_t l, int algo) {
  const int64_t lp = lpad(l);
  const bool x3 = is_x3(algo);
  const int64_t Tpad = round_up(T, TC_KC);
  // project_S: wsum | chunk flags | W image hi | W image lo
  const int64_t bs = align256(lp * 4) + align256(Tpad / TC_KC) + (x3 ? 2 : 1) * align256(lp * Tpad * 4) +
                     align256((round_up(T, 64) + 64) * 4);  // + row_delta | base_nan of the fused statistics pass
  // project_T: rvec | pivot_pad | dscale_pad | partials | Y image hi | Y image lo
  // (the finest split has the largest partial buffer)
  int64_t part = 0;
  for (int kb = 1; kb <= 4; kb *= 2) {
    const TGeom g = t_geometry(T, S, t_cap(algo), kb, 4);
    const int64_t b = (int64_t)g.splits * g.rows_pad * lp * 4;
    if (b > part) part = b;
  }
  // (the fp16 passes keep amax | gs | igs | rvec, partials, the fp16 image and the shares of the rank-1 term here)
  const int64_t Spad = round_up(S, 256);
  const int64_t bt = 4 * align256(lp * 4) + 2 * align256(Spad * 4) + align256(part) + (x3 ? 2 : 1) * align256(lp * Spad * 4) +
                     align256(2 * (int64_t)num_sms() * lp * 8);
  return (bs > bt ? bs : bt) + 256;
}

// One launch, as a grid of clusters of p.cl CTAs along x (grid.x is a multiple of p.cl)
typedef void (*TcKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcParams);
static int launch_kernel(TcKernel kern, int threads, const CUtensorMap& mx, const CUtensorMap& mh, const CUtensorMap& ml,
                         const TcParams& p, dim3 grid, size_t smem, cudaStream_t stream) {
  XB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (p.cl <= 1) {
    kern<<<grid, threads, smem, stream>>>(mx, mh, ml, p);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)p.cl;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    XB_CUDA(cudaLaunchKernelEx(&cfg, kern, mx, mh, ml, p));
  }
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

template <int NS, bool SIDE_T, int KB, int RN = 0, bool PK = false, bool TFAST = false, int MODE = MODE_F32>
static int launch_tc(const CUtensorMap& mx, const CUtensorMap& mh, const CUtensorMap& ml, const TcParams& p, dim3 grid,
                     size_t smem, cudaStream_t stream) {
  return launch_kernel(project_tc_kernel<NS, SIDE_T, KB, false, RN, PK, TFAST, MODE>, tc_threads(tc_wide(NS, SIDE_T, RN != 0)), mx,
                       mh, ml, p, grid, smem, stream);
}
// project_T, single product, with MODE (the pass that writes the fp16 copy / the passes that read it)
template <int MODE>
static int launch_T_mode(int kb, bool fast, const CUtensorMap& mx, const CUtensorMap& mh, const TcParams& p, dim3 grid,
                         size_t smem, cudaStream_t stream) {
  if (fast || MODE == MODE_H16)
    return kb == 1   ? launch_tc<1, true, 1, 0, false, true, MODE>(mx, mh, mh, p, grid, smem, stream)
           : kb == 2 ? launch_tc<1, true, 2, 0, false, true, MODE>(mx, mh, mh, p, grid, smem, stream)
                     : launch_tc<1, true, 4, 0, false, true, MODE>(mx, mh, mh, p, grid, smem, stream);
  return kb == 1   ? launch_tc<1, true, 1, 0, false, false, MODE>(mx, mh, mh, p, grid, smem, stream)
         : kb == 2 ? launch_tc<1, true, 2, 0, false, false, MODE>(mx, mh, mh, p, grid, smem, stream)
                   : launch_tc<1, true, 4, 0, false, false, MODE>(mx, mh, mh, p, grid, smem, stream);
}
// project_T: pick the instantiation for (KB, no-NaN promise)
template <int NS, int RN, bool PK>
static int launch_T(int kb, bool fast, const CUtensorMap& mx, const CUtensorMap& mh, const CUtensorMap& ml, const TcParams& p,
                    dim3 grid, size_t smem, cudaStream_t stream) {
  if (fast)
    return kb == 1   ? launch_tc<NS, true, 1, RN, PK, true>(mx, mh, ml, p, grid, smem, stream)
           : kb == 2 ? launch_tc<NS, true, 2, RN, PK, true>(mx, mh, ml, p, grid, smem, stream)
                     : launch_tc<NS, true, 4, RN, PK, true>(mx, mh, ml, p, grid, smem, stream);
  return kb == 1   ? launch_tc<NS, true, 1, RN, PK, false>(mx, mh, ml, p, grid, smem, stream)
         : kb == 2 ? launch_tc<NS, true, 2, RN, PK, false>(mx, mh, ml, p, grid, smem, stream)
                   : launch_tc<NS, true, 4, RN, PK, false>(mx, mh, ml, p, grid, smem, stream);
}

int project_S_tc(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot, const float* dscale,
                 const float* ccorr, const uint8_t* row_valid, const float* W, int64_t ldw, int64_t l, float* Yt,
                 int64_t ldy, void* workspace, int64_t workspace_bytes, int algo, cudaStream_t stream) {
  (void)workspace_bytes;
  const int lp = (int)lpad(l);
  const int ns = algo_ns(algo);
  const int64_t Tpad = round_up(T, TC_KC);
  uint8_t* ws = (uint8_t*)workspace;
  float* wsum = (float*)ws;
  uint8_t* flags = ws + align256(lp * 4);
  const bool pk = use_pack(ns, lp);
  float* Whi = (float*)(flags + align256(Tpad / TC_KC));
  // packed: one image per slab, [hi rows | lo rows]
  float* Wlo = ns == 3 ? (pk ? Whi + lp * 32 : (float*)((uint8_t*)Whi + align256(lp * Tpad * 4))) : nullptr;
  int rc = XEOFS_OK;
  if (ccorr) {  // the rank-1 term ccorr[s] * colsum(W)[j] exists only for un-centred fields
    rc = launch_colsum(W, T, ldw, lp, row_valid, wsum, stream);
    if (rc) return rc;
  }
  if (row_valid) {
    chunk_flags_kernel<<<(unsigned)ceil_div(Tpad / TC_KC, 128), 128, 0, stream>>>(row_valid, T, (int)(Tpad / TC_KC), flags);
    XB_LAUNCH_CHECK();
  }
  prep_W_kernel<<<(unsigned)(Tpad / TC_KC), 256, 0, stream>>>(W, T, ldw, lp, Whi, Wlo, algo_rn(algo) == 1 ? 1 : 0,
                                                              pk ? 2 * lp * 32 : 0);
  XB_LAUNCH_CHECK();
  CUtensorMap mx, mh, ml;
  rc = make_map2(&mx, X, S, T, ldx, TC_TILE, TC_KC, false);
  if (rc) return rc;
  // the operand images as rows of 1 KB: a stage's image arrives as lp/8 long pieces
  const int bp = pk ? 2 : 1;
  rc = make_map2(&mh, Whi, 256, (Tpad / TC_KC) * (bp * lp / 8), 256, 256, bp * lp / 8, false);
  if (rc) return rc;
  ml = mh;
  if (ns == 3 && !pk) {
    rc = make_map2(&ml, Wlo, 256, (Tpad / TC_KC) * (lp / 8), 256, 256, lp / 8, false);
    if (rc) return rc;
  }
  const Shape sh = pick_shape(lp, ns, false, 1, algo_rn(algo) != 0);
  TcParams p{};
  p.T = T; p.S = S; p.lp = lp;
  p.stages = sh.stages; p.dcols = sh.dcols; p.tmem_cols = sh.tmem_cols;
  p.nchunks_total = (int)(Tpad / TC_KC);
  p.chunks_per_cta = p.nchunks_total;
  p.pivot = pivot; p.dscale = dscale; p.ccorr = ccorr; p.wsum = wsum;
  p.out = Yt; p.ldo = ldy;
  p.X = X; p.ldx = ldx; p.bimg_hi = Whi; p.bimg_lo = Wlo;
  p.b_bulk = env_int("XEOFS_TC_BBULK", 1);
  p.cl = pick_cluster();
  p.chunk_flags = row_valid ? flags : nullptr;
  dim3 grid((unsigned)round_up(ceil_div(S, TC_TILE), p.cl));
  return ns == 3 && pk  ? launch_tc<3, false, 1, 0, true>(mx, mh, ml, p, grid, sh.smem, stream)
         : ns == 3      ? launch_tc<3, false, 1>(mx, mh, ml, p, grid, sh.smem, stream)
         : ns == 2      ? launch_tc<2, false, 1>(mx, mh, ml, p, grid, sh.smem, stream)
         : algo_rn(algo) == 1 ? launch_tc<1, false, 1, 1>(mx, mh, ml, p, grid, sh.smem, stream)
         : algo_rn(algo) == 2 ? launch_tc<1, false, 1, 2>(mx, mh, ml, p, grid, sh.smem, stream)
                        : launch_tc<1, false, 1>(mx, mh, ml, p, grid, sh.smem, stream);
}

__global__ void stats_init_kernel(double* scalars, int32_t* row_delta, int64_t T, int32_t* base_nan) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < T) row_delta[i] = 0;
  if (i == 0) {
    scalars[0] = 0.0; scalars[1] = 0.0; scalars[2] = 0.0; scalars[3] = 2147483647.0;
    *base_nan = 0;
  }
}
__global__ void row_nan_kernel(const int32_t* __restrict__ row_delta, const int32_t* __restrict__ base_nan, int64_t T,
                               int32_t* __restrict__ row_nan) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < T) row_nan[i] = *base_nan + row_delta[i];
}

// Statistics + first product in one read of X (every sample taken as present: the caller re-does the product if the
// statistics then name samples that are NaN throughout).
int project_S_stats_tc(const float* X, int64_t T, int64_t S, int64_t ldx, const double* featw, int flags, const float* W,
                       int64_t ldw, int64_t l, float* mean, float* stdv, uint8_t* valid, float* pivot, float* dscale,
                       float* ccorr, double* scalars, int32_t* row_nan, float* Yt, int64_t ldy, void* workspace,
                       cudaStream_t stream, uint16_t* copy16, int64_t ldc16, float* c0, float* ic16, float* cc16) {
  const int lp = (int)lpad(l);
  const int64_t Tpad = round_up(T, TC_KC);
  uint8_t* ws = (uint8_t*)workspace;
  float* wsum = (float*)ws;
  uint8_t* flags_unused = ws + align256(lp * 4);
  float* Whi = (float*)(flags_unused + align256(Tpad / TC_KC));
  // row_delta | base_nan live behind the operand image
  int32_t* row_delta = (int32_t*)((uint8_t*)Whi + align256(lp * Tpad * 4));
  int32_t* base_nan = row_delta + round_up(T, 64);
  int rc = launch_colsum(W, T, ldw, lp, nullptr, wsum, stream);
  if (rc) return rc;
  prep_W_kernel<<<(unsigned)(Tpad / TC_KC), 256, 0, stream>>>(W, T, ldw, lp, Whi, nullptr);
  XB_LAUNCH_CHECK();
  stats_init_kernel<<<(unsigned)ceil_div(T, 256), 256, 0, stream>>>(scalars, row_delta, T, base_nan);
  XB_LAUNCH_CHECK();
  CUtensorMap mx, mh;
  rc = make_map2(&mx, X, S, T, ldx, TC_TILE, TC_KC, false);
  if (rc) return rc;
  rc = make_map2(&mh, Whi, 256, (Tpad / TC_KC) * (lp / 8), 256, 256, lp / 8, false);
  if (rc) return rc;
  const Shape sh = pick_shape(lp, 1, false, 1);
  TcParams p{};
  p.T = T; p.S = S; p.lp = lp;
  p.stages = sh.stages; p.dcols = sh.dcols; p.tmem_cols = sh.tmem_cols;
  p.nchunks_total = (int)(Tpad / TC_KC);
  p.chunks_per_cta = p.nchunks_total;
  p.wsum = wsum;
  p.out = Yt; p.ldo = ldy;
  p.X = X; p.ldx = ldx; p.bimg_hi = Whi;
  p.b_bulk = env_int("XEOFS_TC_BBULK", 1);
  p.cl = pick_cluster();
  p.featw = featw; p.stat_flags = flags;
  p.mean_out = mean; p.std_out = stdv; p.valid_out = valid; p.pivot_out = pivot; p.dscale_out = dscale; p.ccorr_out = ccorr;
  p.scalars_out = scalars; p.row_delta = row_delta; p.base_nan = base_nan;
  dim3 grid((unsigned)round_up(ceil_div(S, TC_TILE), p.cl));
  if (copy16) {
    // the pass also writes the fp16 copy of the (shifted) field: per-feature power of two from a pre-sample first
    h16_prescale_kernel<<<(unsigned)ceil_div(S, 256), 256, 0, stream>>>(X, T, S, ldx, c0);
    XB_LAUNCH_CHECK();
    p.copy16 = copy16; p.ldc16 = ldc16; p.c0 = c0; p.ic16_out = ic16; p.cc16_out = cc16;
    rc = launch_kernel(project_tc_kernel<1, false, 1, true, 0, false, false, MODE_WCOPY>, tc_threads(false), mx, mh, mh, p, grid,
                       sh.smem, stream);
  } else {
    rc = launch_kernel(project_tc_kernel<1, false, 1, true>, tc_threads(false), mx, mh, mh, p, grid, sh.smem, stream);
  }
  if (rc) return rc;
  row_nan_kernel<<<(unsigned)ceil_div(T, 256), 256, 0, stream>>>(row_delta, base_nan, T, row_nan);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

int project_T_tc(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot, const float* dscale,
                 const float* ccorr, const uint8_t* row_valid, const float* Yt, int64_t ldy, int64_t l, float* Z,
                 int64_t ldz, void* workspace, int64_t workspace_bytes, int algo, bool no_nan, cudaStream_t stream,
                 uint16_t* copy16, int64_t ldc16, const float* e16) {
  (void)workspace_bytes;
  const int lp = (int)lpad(l);
  const int ns = algo_ns(algo);
  const Shape sh = pick_shape_T(lp, ns, S, ldx, algo_rn(algo) != 0, copy16 != nullptr);
  XB_CHECK_ARG(sh.stages >= 1, "project_T: no pipeline shape fits lp=%d", lp);
  const int kb = sh.kb;
  const int cl = pick_cluster();
  const TGeom g = t_geometry(T, S, t_cap(algo), kb, cl);
  XB_CHECK_ARG(g.splits <= 65535, "project_T: too many splits");
  uint8_t* ws = (uint8_t*)workspace;
  float* rvec = (float*)ws; ws += align256(lp * 4);
  float* pdpad = (float*)ws; ws += 2 * align256(g.Spad * 4);
  float* part = (float*)ws; ws += align256((int64_t)g.splits * g.rows_pad * lp * 4);
  const bool pk = use_pack(ns, lp);
  float* Yhi = (float*)ws; ws += align256(lp * g.Spad * 4);
  float* Ylo = ns == 3 ? (pk ? Yhi + lp * 32 : (float*)ws) : nullptr;
  // (the pass that writes the fp16 copy carries e16 where dscale would go: dscale itself rides on the operand image)
  pad_vectors_kernel<<<(unsigned)ceil_div(g.Spad, 256), 256, 0, stream>>>(pivot, copy16 ? e16 : dscale, S, g.Spad, pdpad);
  XB_LAUNCH_CHECK();
  tile_Y_kernel<<<(unsigned)(g.Spad / TC_KC), 256, 0, stream>>>(Yt, S, ldy, lp, Yhi, Ylo, algo_rn(algo) == 1 ? 1 : 0,
                                                                pk ? 2 * lp * 32 : 0, (ns != 2 && algo_rn(algo) != 2) ? dscale : nullptr);
  XB_LAUNCH_CHECK();
  int rc;
  if (ccorr) {
    rc = launch_ccorr_dot(Yt, S, ldy, ccorr, lp, rvec, stream);
    if (rc) return rc;
  }
  TcParams p{};
  p.T = T; p.S = S; p.lp = lp;
  p.stages = sh.stages; p.dcols = sh.dcols; p.tmem_cols = sh.tmem_cols;
  p.nchunks_total = g.chunks_total;
  p.chunks_per_cta = g.chunks_per_cta;
  p.pivot = pdpad; p.dscale = nullptr; p.ccorr = nullptr; p.wsum = nullptr;
  p.out = part; p.ldo = lp;
  p.X = X; p.ldx = ldx; p.bimg_hi = Yhi; p.bimg_lo = Ylo;
  p.b_bulk = env_int("XEOFS_TC_BBULK", 1);
  p.cl = cl;
  dim3 grid((unsigned)g.t_tiles, (unsigned)g.splits);
  CUtensorMap mx, mh, ml;
  // rows of X in pieces of KB*32 + 4 floats (the 4 extra only give the shared-memory rows their odd pitch)
  rc = make_map2(&mx, X, S, T, ldx, kb * TC_KC + 4, TC_TILE, false);
  if (rc) return rc;
  const int bp = pk ? 2 : 1;
  rc = make_map2(&mh, Yhi, 256, (g.Spad / TC_KC) * (bp * lp / 8), 256, 256, kb * bp * lp / 8, false);
  if (rc) return rc;
  ml = mh;
  if (ns == 3 && !pk) {
    rc = make_map2(&ml, Ylo, 256, (g.Spad / TC_KC) * (lp / 8), 256, 256, kb * lp / 8, false);
    if (rc) return rc;
  }
  const bool fast = no_nan && !row_valid;
  p.copy16 = copy16; p.ldc16 = ldc16;
  if (copy16) {
    XB_CHECK_ARG(ns == 1 && algo_rn(algo) == 0 && !ccorr && !row_valid && e16 && ldc16 >= g.Spad && ldc16 % 8 == 0,
                 "project_T: the fp16 copy is written by a single-TF32 pass over a centred field with every sample present");
    rc = launch_T_mode<MODE_WCOPY>(kb, fast, mx, mh, p, grid, sh.smem, stream);
  } else if (ns == 3 && pk) rc = launch_T<3, 0, true>(kb, fast, mx, mh, ml, p, grid, sh.smem, stream);
  else if (ns == 3) rc = launch_T<3, 0, false>(kb, fast, mx, mh, ml, p, grid, sh.smem, stream);
  else if (ns == 2) rc = launch_T<2, 0, false>(kb, fast, mx, mh, ml, p, grid, sh.smem, stream);
  else if (algo_rn(algo) == 1) rc = launch_T<1, 1, false>(kb, fast, mx, mh, ml, p, grid, sh.smem, stream);
  else if (algo_rn(algo) == 2) rc = launch_T<1, 2, false>(kb, true, mx, mh, ml, p, grid, sh.smem, stream);
  else rc = launch_T<1, 0, false>(kb, fast, mx, mh, ml, p, grid, sh.smem, stream);
  if (rc) return rc;
  reduce_partials_kernel<<<(unsigned)ceil_div(T * lp, 256), 256, 0, stream>>>(part, g.splits, g.rows_pad, lp, T,
                                                                              ccorr ? rvec : nullptr, row_valid, Z, ldz);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

// ------------------------------------------------------------------------------------------------ the fp16 passes
// project_S / project_T on the fp16 copy A16 (T x S, pitch ldc halves) written by the MODE_WCOPY pass:
// A[t,s] = A16[t,s] ic16[s].  Single product (kind::f16); the small operand goes into its image as fp16 with a power
// of two per column (from its absolute maximum) that the epilogue takes out again.
int h16_scales(const float* dscale, const float* stdv, int64_t S, float* e16, float* ic16, cudaStream_t stream) {
  h16_scales_kernel<<<(unsigned)ceil_div(S, 256), 256, 0, stream>>>(dscale, stdv, S, e16, ic16);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

int project_S16_tc(const uint16_t* A16, int64_t T, int64_t S, int64_t ldc, const float* ic16, const float* cc16, const float* W,
                   int64_t ldw, int64_t l, float* Yt, int64_t ldy, void* workspace, cudaStream_t stream) {
  const int lp = (int)lpad(l);
  const int64_t Tpad = round_up(T, 64);
  uint8_t* ws = (uint8_t*)workspace;
  float* amax = (float*)ws; ws += align256(lp * 4);
  float* gs = (float*)ws; ws += align256(lp * 4);
  float* igs = (float*)ws; ws += align256(lp * 4);
  float* wsum = (float*)ws; ws += align256(lp * 4);
  uint16_t* Wimg = (uint16_t*)ws;
  if (cc16) {  // the copy is shifted, not centred: rank-1 term cc16[s] * colsum(W)[j]
    int rc0 = launch_colsum(W, T, ldw, lp, nullptr, wsum, stream);
    if (rc0) return rc0;
  }
  XB_CUDA(cudaMemsetAsync(amax, 0, lp * 4, stream));
  absmax_cols_kernel<<<(unsigned)imin(T, 512), 128, 0, stream>>>(W, T, ldw, lp, 0, nullptr, amax, nullptr, nullptr);
  h16_colscale_kernel<<<1, 128, 0, stream>>>(amax, lp, gs, nullptr, 0, nullptr);
  recip_kernel<<<1, 128, 0, stream>>>(gs, lp, igs);
  prep_W16_kernel<<<(unsigned)(Tpad / 64), 256, 0, stream>>>(W, T, ldw, lp, gs, Wimg);
  XB_LAUNCH_CHECK();
  CUtensorMap mx;
  int rc = make_map16(&mx, A16, S, T, ldc, TC_TILE, 64);
  if (rc) return rc;
  const Shape sh = pick_shape(lp, 1, false, 1);
  TcParams p{};
  p.T = T; p.S = S; p.lp = lp;
  p.stages = sh.stages; p.dcols = sh.dcols; p.tmem_cols = sh.tmem_cols;
  p.nchunks_total = (int)(Tpad / 64);
  p.chunks_per_cta = p.nchunks_total;
  p.dscale = ic16; p.igs = igs; p.ccorr = cc16; p.wsum = wsum;
  p.out = Yt; p.ldo = ldy;
  p.bimg_hi = (const float*)Wimg;
  p.b_bulk = 1;
  p.cl = 1;
  dim3 grid((unsigned)ceil_div(S, TC_TILE));
  return launch_tc<1, false, 1, 0, false, false, MODE_H16>(mx, mx, mx, p, grid, sh.smem, stream);
}

int project_T16_tc(const uint16_t* A16, int64_t T, int64_t S, int64_t ldc, const float* ic16, const float* cc16, const float* Yt,
                   int64_t ldy, int64_t l, float* Z, int64_t ldz, void* workspace, cudaStream_t stream) {
  const int lp = (int)lpad(l);
  const Shape sh = pick_shape_T(lp, 1, S, ldc);
  XB_CHECK_ARG(sh.stages >= 1, "project_T16: no pipeline shape fits lp=%d", lp);
  const int kb = sh.kb;
  const TGeom g = t_geometry(T, S, 1024, kb, 1, 64);
  XB_CHECK_ARG(g.splits <= 65535, "project_T16: too many splits");
  const int64_t Spad = round_up(S, 256);
  uint8_t* ws = (uint8_t*)workspace;
  float* amax = (float*)ws; ws += align256(lp * 4);
  float* gs = (float*)ws; ws += align256(lp * 4);
  float* igs = (float*)ws; ws += align256(lp * 4);
  float* rvec = (float*)ws; ws += align256(lp * 4);
  float* part = (float*)ws; ws += align256((int64_t)g.splits * g.rows_pad * lp * 4);
  uint16_t* Yimg = (uint16_t*)ws;
  double* dpart = (double*)(ws + align256(lp * Spad * 2));  // inside the fp32-image budget of tc_workspace_bytes
  const int nparts = (int)imin(ceil_div(S, 256 * 8), 2 * (int64_t)num_sms());
  XB_CUDA(cudaMemsetAsync(amax, 0, lp * 4, stream));
  // one read of Yt: column maxima and the rank-1 term of the shifted copy, r[j] = sum_s cc16[s] Yt[j,s]
  absmax_cols_kernel<<<dim3((unsigned)nparts, (unsigned)lp), 256, 0, stream>>>(Yt, S, ldy, lp, 1, ic16, amax, cc16,
                                                                              cc16 ? dpart : nullptr);
  h16_colscale_kernel<<<1, 128, 0, stream>>>(amax, lp, gs, cc16 ? dpart : nullptr, nparts, rvec);
  recip_kernel<<<1, 128, 0, stream>>>(gs, lp, igs);
  tile_Y16_kernel<<<(unsigned)(Spad / 64), 256, 0, stream>>>(Yt, S, ldy, lp, ic16, gs, Yimg);
  XB_LAUNCH_CHECK();
  CUtensorMap mx;
  int rc = make_map16(&mx, A16, S, T, ldc, kb * 64 + 8, TC_TILE);
  if (rc) return rc;
  TcParams p{};
  p.T = T; p.S = S; p.lp = lp;
  p.stages = sh.stages; p.dcols = sh.dcols; p.tmem_cols = sh.tmem_cols;
  p.nchunks_total = g.chunks_total;
  p.chunks_per_cta = g.chunks_per_cta;
  p.out = part; p.ldo = lp;
  p.bimg_hi = (const float*)Yimg;
  p.b_bulk = 1;
  p.cl = 1;
  dim3 grid((unsigned)g.t_tiles, (unsigned)g.splits);
  rc = launch_T_mode<MODE_H16>(kb, true, mx, mx, p, grid, sh.smem, stream);
  if (rc) return rc;
  reduce_partials_kernel<<<(unsigned)ceil_div(T * lp, 256), 256, 0, stream>>>(part, g.splits, g.rows_pad, lp, T,
                                                                              cc16 ? rvec : nullptr, nullptr, Z, ldz, igs);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

}  // namespace xb
