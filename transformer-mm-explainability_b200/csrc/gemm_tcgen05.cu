// tcgen05 (5th-gen tensor core) GEMM with fp32-faithful numerics:  C[M,N] = A[M,K] * Bt[N,K]^T  (+ epilogue).
//
// Why 3xTF32: the relevancy maps must match the fp32 reference to 1e-4 after ~100 chained GEMMs (forward +
// dgrad), which rules out plain TF32/BF16 inputs.  Each fp32 operand x is split exactly into
//      x = hi + lo,   hi = x with the low 13 mantissa bits cleared (a TF32 value),  lo = x - hi (exact in fp32)
// and the product is A_hi*B_hi + (A_lo*B_hi + A_hi*B_lo); the dropped lo*lo term and the TF32 rounding of lo are both
// ~2^-21 relative.  The tensor core adds into its fp32 accumulator with truncation (measured: error grows linearly
// with the number of accumulating MMAs), so the two small cross products go to their OWN TMEM accumulator (their
// truncation error is 2^-11 smaller) and are added to the hi*hi accumulator once, in the epilogue, with RN.
//
// A 3-pass fp32 GEMM is limited by operand movement, not by the tensor pipe: measured on B200, L2->SM delivers
// ~24 B/clk/SM with all 148 SMs pulling, and shared memory serves 128 B/clk/SM to TMA writes, LSU traffic and the
// UMMA operand fetch together.  So both operands cross L2 ONCE as raw fp32 and are split on the SM:
//   * A (activations): raw tile -> registers -> hi/lo -> TMEM (tcgen05.st); the MMAs take A from TMEM (.ts form),
//     so A costs one shared-memory read instead of three operand fetches;
//   * B (weights): raw tile split in shared memory (hi in place, lo beside it), fetched by the MMAs from there.
//
// Tile shape: 128 x BN with BN in {128, 144, 160}, chosen per launch so that the tile count fills whole waves of the
// 148 SMs (the batch-64 vision tower has 25 row tiles: 128-wide tiles give 150 / 450 / 600 tiles = 2 / 4 / 5 waves for
// 1.01 / 3.04 / 4.05 waves of work).  Every output element sums its K products in the same order whatever BN is, so
// the choice never changes a result bit.  TMEM holds 2*BN accumulator columns + 64 columns of A per stage, which is
// why the wide tiles run a 3-stage ring.
//
// Structure (one CTA per SM, persistent over 128xBN output tiles, 3/4-stage ring of 128x32 K-slabs):
//   warp 0      TMA producer: raw fp32 A and B tiles -> shared memory (128-byte swizzle), mbarrier complete_tx
//   warps 12-15 splitters (one per TMEM lane quarter; build with -DMMX_TC_SPLIT_WARPS=8 for two per quarter): A tile (smem) -> hi/lo -> TMEM columns of this
//               stage; B tile -> hi/lo planes in smem
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma.kind::tf32 (3 per k-step), tcgen05.commit
//   warp 2      TMEM allocator (512 columns: 128 main + 128 cross accumulator + 4 stages x 64 columns of A)
//   warps 4-11  epilogue: tcgen05.ld (main + cross, added with RN) -> registers, TMEM released at once, then
//               bias / act' / residual / act -> global from registers (overlaps the next tile's main loop)
#include "gemm.cuh"
#include "gemm_epilogue.cuh"
#include <mutex>
#include "tcgen05_ptx.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdlib>

namespace mmx {

namespace tc {

constexpr int BM = 128, BK = 32;              // BK fp32 = 128 bytes = one swizzle-128B row; BN is a template parameter
// Splitter warps: 4 (one per TMEM lane quarter, 512 threads, 128 regs/thread for everyone) or 8 (two per quarter, each
// converting half of the slab's K columns; 640 threads, which needs setmaxnreg rebalancing because the epilogue uses
// ~128 registers: 4*32*64 + 8*32*136 + 8*32*72 = 61440 = 640*96).  Measured 75.7 vs ~77 us on the QKV shape; the
// 4-warp build is the default because it has no blocking register hand-off.
#ifndef MMX_TC_SPLIT_WARPS
#define MMX_TC_SPLIT_WARPS 4
#endif
constexpr int SPLIT_WARPS = MMX_TC_SPLIT_WARPS;
constexpr bool REBALANCE = SPLIT_WARPS == 8;
constexpr int THREADS = (12 + SPLIT_WARPS) * 32;
constexpr int REGS_CTRL = 64, REGS_EPI = 136, REGS_SPLIT = 72;
constexpr int EPI_WARP0 = 4, EPI_WARPS = 8, SPLIT_WARP0 = 12;
constexpr int A_BYTES = BM * BK * 4;           // 16 KB raw A slab
template <int BN, int STAGES> struct Cfg {
  static_assert(BN % 16 == 0 && BN >= 128 && BN <= 160, "UMMA N for M=128: multiple of 16; TMEM budget caps it at 160");
  static constexpr int B_BYTES = BN * BK * 4;          // per B plane (raw / lo): 16-20 KB, a multiple of 1024
  static constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;
  static constexpr int EPI_STG = EPI_WARPS * EPI_STAGE_FLOATS * 4;     // epilogue staging: 16 KB
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_STG;
  // TMEM column map: main accumulator, cross accumulator, then the A ring (stage s: hi at TM_A+64s, lo at +32)
  static constexpr uint32_t TM_MAIN = 0, TM_CROSS = (BN + 31) / 32 * 32, TM_A = 2 * TM_CROSS;
  static_assert(TM_A + 64 * STAGES <= 512, "TMEM has 512 columns");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory per CTA");
  static constexpr int CW = BN / 2;                     // accumulator columns per epilogue warp (64 / 72 / 80)
};

using namespace ::mmx::tcp;

#ifdef MMX_TC_TRACE   // build with -DMMX_TC_TRACE for profiles/gemm_trace.py; off in the product (it costs ~15 %)
#define MMX_TRACE(role, idx, slot)                                                                  \
  do {                                                                                               \
    if (p.trace != nullptr && blockIdx.x == 0 && (idx) < 256) p.trace[((role) * 256 + (idx)) * 4 + (slot)] = clock64(); \
  } while (0)
#else
#define MMX_TRACE(role, idx, slot) do { } while (0)
#endif

struct Params {
  int M, N, K, ldc;
  long long* trace;   // optional device buffer [4 roles][256 events][4]: clock64 timeline of CTA 0 (profiling aid)
  int dbg;   // timing experiments only (results invalid): 1 = no MMAs, 2 = no splitter work, 4 = no TMA, 8 = no epilogue stores
  float* C;
  GemmEpilogue ep;
  // batched form (the relevancy updates: one launch for all samples): tile t belongs to sample t / (tiles_m * tiles_n); the
  // operand maps cover the samples stacked along their row dimension (rows_a / rows_b rows per sample), C and the residual
  // advance by stride_c / stride_res elements per sample.  batch = 1: the plain GEMM.
  int batch = 1, rows_a = 0, rows_b = 0;
  long long stride_c = 0, stride_res = 0;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(THREADS, 1) gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                 const __grid_constant__ CUtensorMap mapB, Params p) {
  using cfg = Cfg<BN, STAGES>;
  constexpr int B_BYTES = cfg::B_BYTES, STAGE_BYTES = cfg::STAGE_BYTES, CW = cfg::CW;
  constexpr uint32_t TM_MAIN = cfg::TM_MAIN, TM_CROSS = cfg::TM_CROSS, TM_A = cfg::TM_A;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // swizzle-128B tiles need 1024 B alignment
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  // barriers (8 B each): full_tma[S], split_done[S], empty[S], tmem_full, tmem_empty, then the TMEM base slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto split_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (3 * STAGES), tempty_bar = bar_base + 8u * (3 * STAGES + 1);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * STAGE_BYTES + 8 * (3 * STAGES + 2));
  float* stg_base = reinterpret_cast<float*>(smem_gen + STAGES * STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.ep.m_dev) p.M = min(p.M, __ldg(p.ep.m_dev));       // ragged batch: the host-side M is only an upper bound
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  const int tiles_per = tiles_m * tiles_n;
  const int num_tiles = tiles_per * p.batch;
  const int nk = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(split_bar(s), SPLIT_WARPS); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, EPI_WARPS);
    // (split_done gets one arrival per splitter warp)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (REBALANCE && warp < 4) {
    reg_dec<REGS_CTRL>();
  }
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The whole warp runs the loop (warp-uniform control flow keeps addresses in uniform registers); one elected
    // lane issues.  Single-lane loops made ptxas wrap every UTMALDG / UTCHMMA in a lane "waterfall" loop.
    int stage = 0; uint32_t phase = 0;
    int pslab = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int bi = t / tiles_per, tr = t - bi * tiles_per;
      const int m0 = (tr % tiles_m) * BM, n0 = (tr / tiles_m) * BN;
      const int ra = bi * p.rows_a + m0, rb = bi * p.rows_b + n0;   // rows in the stacked operand maps
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        MMX_TRACE(0, pslab, 0);
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        if (elect_one()) {
          if (p.dbg & 4) { mbar_arrive(full_bar(stage)); }
          else {
            mbar_expect_tx(full_bar(stage), A_BYTES + B_BYTES);
            tma_load_2d(sa, &mapA, full_bar(stage), kb * BK, ra);
            tma_load_2d(sa + A_BYTES, &mapB, full_bar(stage), kb * BK, rb);
          }
        }
        __syncwarp();
        MMX_TRACE(0, pslab, 1);
        ++pslab;
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp loops, one lane issues)
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    int stage = 0; uint32_t phase = 0;
    int it = 0, mslab = 0;
    const uint32_t d_main = tmem_base + TM_MAIN, d_cross = tmem_base + TM_CROSS;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      mbar_wait(tempty_bar, (uint32_t)(it & 1) ^ 1);               // epilogue drained the accumulators
      tc_fence_after();
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(split_bar(stage), phase);                       // A hi/lo in TMEM, B lo plane written (implies the TMA landed)
        MMX_TRACE(1, mslab, 0);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        const uint64_t b_hi = make_desc(sa + A_BYTES), b_lo = make_desc(sa + A_BYTES + B_BYTES);
        const uint32_t a_hi = tmem_base + TM_A + 64u * stage, a_lo = a_hi + 32u;
        if (elect_one()) {
          if (!(p.dbg & 1)) {
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {                    // UMMA_K = 8 (tf32): +32 B in smem (+2 in the descriptor), +8 TMEM columns
              const uint64_t adv = (uint64_t)(2 * k);
              const uint32_t first = (kb > 0 || k > 0) ? 1u : 0u;
              umma_tf32_ts(d_cross, a_lo + 8u * k, b_hi + adv, idesc, first);
              umma_tf32_ts(d_cross, a_hi + 8u * k, b_lo + adv, idesc, 1u);
              umma_tf32_ts(d_main, a_hi + 8u * k, b_hi + adv, idesc, first);
            }
          }
          MMX_TRACE(1, mslab, 1);
          umma_commit(empty_bar(stage));                          // frees the stage (smem + TMEM A slab) when these MMAs retire
          if (kb == nk - 1) umma_commit(tfull_bar);               // accumulators complete -> epilogue
        }
        __syncwarp();
        MMX_TRACE(1, mslab, 2);
        ++mslab;
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= SPLIT_WARP0) {
    if (REBALANCE) reg_dec<REGS_SPLIT>();
    // ------------------------------------------------------------------ splitter: A slab (smem) -> hi/lo -> TMEM
    const int q = warp & 3;                                     // TMEM lane quarter this warp may write
    constexpr int HSTEP = SPLIT_WARPS / 4;                       // warps per lane quarter (1 or 2)
    const int half0 = (warp - SPLIT_WARP0) >> 2;                // first 16-column half of the slab this warp converts
    const int stid = threadIdx.x - SPLIT_WARP0 * 32;            // 0 .. SPLIT_WARPS*32-1
    const int row = q * 32 + lane;                              // tile row handled by this thread
    const uint32_t row_off = (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
    int stage = 0; uint32_t phase = 0;
    int slab = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      for (int kb = 0; kb < nk; ++kb, ++slab) {
        mbar_wait(full_bar(stage), phase);
        if (warp == SPLIT_WARP0 && lane == 0) MMX_TRACE(2, slab, 0);
        if (!(p.dbg & 2)) {
          // A slab: this thread's row (128 B, swizzled 16-byte chunks) -> hi/lo -> TMEM columns of the stage.
          // The tensor core reads only the top 19 bits of an fp32 operand (kind::tf32 truncates; measured: results
          // are bit-identical with and without an explicit hi plane), so the raw values ARE the hi operand.
          const uint8_t* a_raw = smem_gen + stage * STAGE_BYTES + row_off;
#pragma unroll
          for (int half = half0; half < 2; half += HSTEP) {
            const uint32_t t_hi = tmem_base + ((uint32_t)(q * 32) << 16) + TM_A + 64u * stage + 16u * half;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint32_t chunk = (uint32_t)(half * 4 + c);  // 16-byte chunk = k 4*chunk .. 4*chunk+3
              const uint4 x = *reinterpret_cast<const uint4*>(a_raw + ((chunk ^ sw) << 4));
              const uint32_t xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                hi[c * 4 + e] = xv[e];
                lo[c * 4 + e] = __float_as_uint(__uint_as_float(xv[e]) - __uint_as_float(xv[e] & 0xFFFFE000u));
              }
            }
            tmem_st16(t_hi, hi);
            tmem_st16(t_hi + 32u, lo);
          }
          // B slab: lo plane beside the raw tile (elementwise, so the swizzle does not matter)
          const uint4* braw = reinterpret_cast<const uint4*>(smem_gen + stage * STAGE_BYTES + A_BYTES);
          uint4* blo = reinterpret_cast<uint4*>(smem_gen + stage * STAGE_BYTES + A_BYTES + B_BYTES);
          constexpr int NT = SPLIT_WARPS * 32, NV = B_BYTES / 16, PER = (NV + NT - 1) / NT;   // 16-byte elements per thread
#pragma unroll
          for (int u0 = 0; u0 < PER; u0 += 4) {
            uint4 x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (u0 + u < PER && (NV % NT == 0 || (u0 + u) * NT + stid < NV)) x[u] = braw[(u0 + u) * NT + stid];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (u0 + u >= PER || (NV % NT != 0 && (u0 + u) * NT + stid >= NV)) break;
              uint4 l;
              l.x = __float_as_uint(__uint_as_float(x[u].x) - __uint_as_float(x[u].x & 0xFFFFE000u));
              l.y = __float_as_uint(__uint_as_float(x[u].y) - __uint_as_float(x[u].y & 0xFFFFE000u));
              l.z = __float_as_uint(__uint_as_float(x[u].z) - __uint_as_float(x[u].z & 0xFFFFE000u));
              l.w = __float_as_uint(__uint_as_float(x[u].w) - __uint_as_float(x[u].w & 0xFFFFE000u));
              blo[(u0 + u) * NT + stid] = l;
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to UMMA
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(split_bar(stage));
        if (warp == SPLIT_WARP0 && lane == 0) MMX_TRACE(2, slab, 1);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + EPI_WARPS) {
    if (REBALANCE) reg_inc<REGS_EPI>();
    // ------------------------------------------------------------------ epilogue (8 warps: lane quarter x column half)
    const int q = warp & 3;                                     // TMEM lane quarter this warp may read
    const int ch = (warp - EPI_WARP0) >> 2;                     // column half: CW of the BN accumulator columns
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int bi = t / tiles_per, tr = t - bi * tiles_per;
      const int m0 = (tr % tiles_m) * BM, n0 = (tr / tiles_m) * BN;
      mbar_wait(tfull_bar, (uint32_t)(it & 1));
      if (warp == EPI_WARP0 && lane == 0) MMX_TRACE(3, it, 0);
      tc_fence_after();
      // drain main + cross accumulators into registers (RN add), then hand TMEM back before any global traffic
      uint32_t acc[CW];
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * CW);
#pragma unroll
      for (int c = 0; c + 16 <= CW; c += 16) {                  // 16-column pieces: main straight into acc, cross added
        uint32_t x[16];
        tmem_ld16_nowait(trow + TM_MAIN + c, acc + c);
        tmem_ld16_nowait(trow + TM_CROSS + c, x);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[c + j] = __float_as_uint(__uint_as_float(acc[c + j]) + __uint_as_float(x[j]));
      }
      if constexpr (CW % 16 == 8) {
        constexpr int c = CW - 8;
        uint32_t x[8];
        tmem_ld8_nowait(trow + TM_MAIN + c, acc + c);
        tmem_ld8_nowait(trow + TM_CROSS + c, x);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[c + j] = __float_as_uint(__uint_as_float(acc[c + j]) + __uint_as_float(x[j]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);                  // next tile's MMAs may start
      if (warp == EPI_WARP0 && lane == 0) MMX_TRACE(3, it, 1);
      if (!(p.dbg & 8))      // coalesced through the warp's staging buffer (gemm_epilogue.cuh)
        epilogue_store_rows<CW, true>(acc, stg_base + (warp - EPI_WARP0) * EPI_STAGE_FLOATS, m0 + q * 32, n0 + ch * CW, p.M, p.N, p.C,
                                      p.ldc, bi * p.stride_c, bi * p.stride_res, p.ep, lane);
      if (warp == EPI_WARP0 && lane == 0) MMX_TRACE(3, it, 2);
    }
  }
  // ---------------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
}  // namespace tc
// get (bn < 0) / set the forced tile width; 0 = automatic
int gemm_tc_tile_n(int bn) {
  static std::atomic<int> forced{-1};
  if (forced.load() < 0) { const char* e = getenv("MMX_TC_BN"); const int v = e ? atoi(e) : 0; forced.store(v == 128 || v == 144 || v == 160 ? v : 0); }
  if (bn == 0 || bn == 128 || bn == 144 || bn == 160) forced.store(bn);
  return forced.load();
}
namespace tc {
static long long* g_trace = nullptr;

static int make_map(CUtensorMap* map, const float* base, int rows, int K, int ld, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return 1;
  }
  return 0;
}

// Tile width for this launch: the BN whose tile count costs the fewest SM-waves x BN (ties and near-ties go to the
// narrower tile, which has the deeper ring).  A ragged launch (row count known only on the device) keeps 128.
// MMX_TC_BN=128|144|160 forces one width (profiling).
static int pick_bn(int M, int N, bool ragged) {
  const int forced = gemm_tc_tile_n(-1);
  if (forced) return forced;
  if (ragged) return 128;
  const int sms = sm_count(), tiles_m = cdiv(M, BM);
  int best = 128;
  double best_cost = 1e30;
  for (int bn : {128, 144, 160}) {
    const double cost = (double)cdiv(tiles_m * cdiv(N, bn), sms) * bn * (bn == 128 ? 1.0 : 1.04);
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

template <int BN, int STAGES>
static int launch_bn(const float* A, int lda, const float* Bt, int ldb, int M, int N, int K, const Params& p, cudaStream_t st) {
  using cfg = Cfg<BN, STAGES>;
  CUtensorMap mapA, mapB;
  MMX_TRY(make_map(&mapA, A, p.batch > 1 ? (p.batch - 1) * p.rows_a + M : M, K, lda, BM));
  MMX_TRY(make_map(&mapB, Bt, p.batch > 1 ? (p.batch - 1) * p.rows_b + N : N, K, ldb, BN));
  // per device: the attribute belongs to the (function, device) pair
  static std::atomic<bool> attr_set[MMX_MAX_DEVICES];
  const int dev = current_device();
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    MMX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        cfg::SMEM_BYTES));
    attr_set[dev].store(true, std::memory_order_release);
  }
  const int tiles = cdiv(M, BM) * cdiv(N, BN) * p.batch;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  gemm_tf32x3_kernel<BN, STAGES><<<grid, THREADS, cfg::SMEM_BYTES, st>>>(mapA, mapB, p);
  MMX_LAUNCH_CHECK();
  return 0;
}

static int launch(const float* A, int lda, const float* Bt, int ldb, float* C, int ldc, int M, int N, int K,
                  const GemmEpilogue& ep, cudaStream_t st, int batch = 1, int rows_a = 0, int rows_b = 0, long long stride_c = 0,
                  long long stride_res = 0) {
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("MMX_TC_DBG"); dbg = e ? atoi(e) : 0; }
  Params p{M, N, K, ldc, g_trace, dbg, C, ep};
  p.batch = batch; p.rows_a = rows_a; p.rows_b = rows_b; p.stride_c = stride_c; p.stride_res = stride_res;
  switch (pick_bn(M * batch, N, ep.m_dev != nullptr)) {
    case 144: return launch_bn<144, 3>(A, lda, Bt, ldb, M, N, K, p, st);
    case 160: return launch_bn<160, 3>(A, lda, Bt, ldb, M, N, K, p, st);
    default: return launch_bn<128, 4>(A, lda, Bt, ldb, M, N, K, p, st);
  }
}

}  // namespace tc

int gemm_tc_available() {
  // per device (a process may drive several GPUs); -1 = not probed yet
  static std::atomic<int> avail[MMX_MAX_DEVICES];
  static std::atomic<bool> init{false};
  static std::mutex mu;
  if (!init.load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lk(mu);
    if (!init.load()) {
      for (auto& a : avail) a.store(-1);
      init.store(true, std::memory_order_release);
    }
  }
  const int dev = current_device();
  int a = avail[dev].load(std::memory_order_acquire);
  if (a >= 0) return a;
  std::lock_guard<std::mutex> lk(mu);
  a = 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess && prop.major == 10) {   // tcgen05 / TMEM: sm_100 family only
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn != nullptr) {
      tc::g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
      a = 1;
    }
  }
  avail[dev].store(a, std::memory_order_release);
  return a;
}

// Profiling aid: the next tensor-core GEMM launches write CTA 0's clock64 timeline into `buf`
// ([4 roles][256 events][4] int64: producer, MMA issuer, splitter, epilogue); nullptr switches it off.
void gemm_tc_set_trace(long long* buf) { tc::g_trace = buf; }

// Which problems go to the tensor-core kernel.  Deliberately independent of M, so that a sample computed alone
// and the same sample inside a batch take the same arithmetic path (bitwise-equal maps, sharded == single GPU).
bool gemm_tc_shape_ok(const float* A, int lda, const float* Bt, int ldb, float* C, int ldc, int N, int K,
                      const GemmEpilogue& ep) {
  if (!gemm_tc_available()) return false;
  if (N < 128 || K < 64 || (N % 4) || (lda % 4) || (ldb % 4) || (ldc % 4)) return false;   // K itself is free: TMA zero-fills the tail
  if (!aligned16(A) || !aligned16(Bt) || !aligned16(C)) return false;
  if ((ep.bias && !aligned16(ep.bias)) || (ep.pre && (!aligned16(ep.pre) || ep.ldpre % 4)) ||
      (ep.residual && (!aligned16(ep.residual) || ep.ldres % 4)) || (ep.C_act && !aligned16(ep.C_act)))
    return false;
  return true;
}

int gemm_nt_tc(const float* A, int lda, const float* Bt, int ldb, float* C, int ldc, int M, int N, int K,
               const GemmEpilogue& ep, cudaStream_t st) {
  if (M == 0 || N == 0) return 0;
  return tc::launch(A, lda, Bt, ldb, C, ldc, M, N, K, ep, st);
}

// `batch` independent products in ONE launch: C[b] = A[b] Bt[b]^T (+ residual[b]); A[b] starts rows_a rows after A[b-1] (same
// lda), Bt[b] rows_b rows after Bt[b-1]; a tile's operand boxes may run into the next sample's rows - those rows / columns
// of the product are never stored.
int gemm_nt_tc_batched(const float* A, int lda, int rows_a, const float* Bt, int ldb, int rows_b, float* C, int ldc, long long stride_c,
                       long long stride_res, int M, int N, int K, int batch, const GemmEpilogue& ep, cudaStream_t st) {
  if (M == 0 || N == 0 || batch == 0) return 0;
  MMX_REQUIRE(ep.m_dev == nullptr && ep.pre == nullptr && ep.C_act == nullptr && ep.bias == nullptr, "batched GEMM: residual epilogue only");
  return tc::launch(A, lda, Bt, ldb, C, ldc, M, N, K, ep, st, batch, rows_a, rows_b, stride_c, stride_res);
}

}  // namespace mmx
