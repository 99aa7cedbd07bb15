// tcgen05 GEMM with fp32-faithful numerics from THREE fp16 tensor-core passes:  C[M,N] = A[M,K] * W[N,K]^T (+ epilogue).
//
// Each fp32 operand x is represented as  x = hi + lo' / 2048,  hi = fp16_rn(x),  lo' = fp16_rn((x - hi) * 2048)
// (22 significand bits, like the tf32 hi/lo split of gemm_tcgen05.cu; the 2^11 factor keeps lo' in fp16's normal
// range), and   A W^T = A_hi W_hi^T + (A_lo' W_hi^T + A_hi W_lo'^T) / 2048.   fp16 x fp16 products are exact in the
// fp32 accumulator; the two cross products have their own TMEM accumulator (the tensor core adds with truncation, so
// keeping the 2^-11-smaller terms apart keeps their truncation error 2^-11 smaller) and are folded in once, with RN,
// in the epilogue.  kind::f16 runs at twice the kind::tf32 rate and one instruction covers K = 16 instead of 8, so
// against the tf32x3 kernel this halves the tensor time, the MMA issue count and the per-K-slab barrier round trips.
//
// Range.  fp16 spans 2^-24 .. 65504: a value above 65504 becomes inf (and the result NaN - loud, never silently wrong);
// values below 2^-14 keep an ABSOLUTE precision of 2^-35, so tensors whose typical magnitude is above ~1e-5 lose
// nothing.  Forward activations and weights of the models on this path sit in 1e-3 .. 1e2.  Gradients do not, which is
// why the CLIP engine normalises the seed gradient of every sample to a power of two near 1 and undoes the factor when
// dA is staged (clip_engine.cu / attention.cu) - an exact operation, and a per-sample one, so batch invariance holds.
//
// Operand movement (what bounds a 3-pass GEMM, see profiles/gemm_ablation_r1.md): the weights are static, so their
// hi / lo planes are built ONCE (pack_f16x3, 4 bytes per element like the fp32 original) and arrive by TMA ready for
// the MMAs - no in-kernel B split, no extra shared-memory pass.  A (activations / gradients, produced as fp32 by the
// previous kernel) crosses L2 once as raw fp32, is split in registers and written to TMEM (tcgen05.st, packed fp16
// pairs), and the MMAs take it from there (.ts form).
//
// Structure (one CTA per SM, persistent over 128 x BN tiles, 128 x 64 K-slabs).  The two operands live in SEPARATE rings,
// because they are held for different times: a raw A slab is dead as soon as the splitters have moved it into TMEM, a W slab
// only when its MMAs have retired: A has 2 shared-memory stages (+ 3-4 slabs of split A in TMEM), W has 3-4, and 16 KB
// belong to the epilogue's staging buffers (gemm_epilogue.cuh).  Timeline of one CTA on the QKV product (mmx_gemm_trace,
// profiles/gemm_trace_r2_after.log): 1050 clk per K-slab (the F2FP conversions of the split: 1024 clk of XU work per slab;
// MMA floor 864), ~2000 clk between the last MMA of a tile and the first of the next (one accumulator set: TMEM is full),
// 5-7 k clk of epilogue stores per tile hidden behind the next tile's main loop.
//   warp 0      TMA producer A: raw fp32 (two 32-wide swizzle-128B boxes)            ring: full_a / empty_a   (SA = 2)
//   warp 3      TMA producer W: W_hi, W_lo fp16 (64-wide boxes)                      ring: full_b / empty_b   (SB = 3 or 4)
//   warps 12-15 splitters: one A row per thread -> hi / lo' fp16 pairs -> TMEM slab  ring: split_done / tmem_free (TS = 3 or 4);
//               the shared-memory stage is handed back (empty_a) as soon as the row is in registers
//   warp 1      MMA issuer: 12 tcgen05.mma.kind::f16 per slab (4 k-steps x 3 products); tcgen05.commit -> empty_b, tmem_free
//   warp 2      TMEM allocator (512 columns: BN main + BN cross accumulators + TS x 64 columns of A)
//   warps 4-11  epilogue: tcgen05.ld main + cross/2048 -> registers, TMEM released, then bias / act' / residual / act and
//               the stores, coalesced through a per-warp staging buffer (gemm_epilogue.cuh)
// The kernel is launched as a programmatic dependent: everything before griddepcontrol.wait overlaps the previous kernel.
#include "gemm.cuh"
#include "gemm_epilogue.cuh"
#include "tcgen05_ptx.cuh"
#include <cuda_fp16.h>
#include <cudaTypedefs.h>
#include <cstdlib>

namespace mmx {
namespace hx {

using namespace ::mmx::tcp;

constexpr int BM = 128, BK = 64, SA = 2;      // SA: shared-memory stages of raw A
constexpr int EPI_WARP0 = 4, EPI_WARPS = 8, SPLIT_WARP0 = 12;
// G splitter warps per TMEM lane quarter (4 G in all): each converts 64 / G of a slab's 64 K-columns for its 32 rows
constexpr int threads_for(int G) { return (12 + 4 * G) * 32; }
constexpr int A_SUB = BM * 32 * 4;             // one 128 x 32 fp32 swizzle-128B box: 16 KB
constexpr int A_BYTES = 2 * A_SUB;             // raw A slab 128 x 64 fp32
constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;

#define HX_TRACE(role, idx, slot)                                                                        \
  do {                                                                                                   \
    if (p.trace != nullptr && blockIdx.x == 0 && (idx) < 64) p.trace[((role) * 64 + (idx)) * 4 + (slot)] = clock64(); \
  } while (0)

template <int BN> struct Cfg {
  static_assert(BN % 16 == 0 && BN >= 128 && BN <= 160, "UMMA N for M=128: multiple of 16; TMEM budget caps it at 160");
  static constexpr int B_BYTES = BN * BK * 2;           // per plane (hi / lo): 16-20 KB, a multiple of 1024
  static constexpr int B_STAGE = 2 * B_BYTES;           // hi plane then lo plane
  static constexpr int EPI_STG = EPI_WARPS * EPI_STAGE_FLOATS * 4;      // epilogue staging (gemm_epilogue.cuh): 16 KB
  static constexpr int SB = (227 * 1024 - 1024 - 512 - EPI_STG - SA * A_BYTES) / B_STAGE > 5 ? 5 : (227 * 1024 - 1024 - 512 - EPI_STG - SA * A_BYTES) / B_STAGE;
  static constexpr int SMEM_BYTES = SA * A_BYTES + SB * B_STAGE + 1024 /*align*/ + 512 /*barriers*/ + EPI_STG;
  static constexpr uint32_t TM_MAIN = 0, TM_CROSS = (BN + 31) / 32 * 32, TM_A = 2 * TM_CROSS;   // A slabs: hi at +64s, lo at +32
  static constexpr int TS = (512 - (int)TM_A) / 64;     // TMEM slabs of split A
  static_assert(SB >= 3 && TS >= 2, "ring depths");
  static_assert(TM_A + 64 * TS <= 512, "TMEM has 512 columns");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory per CTA");
  static constexpr int CW = BN / 2;                     // accumulator columns per epilogue warp
};

struct Params {
  int M, N, K, ldc;
  float* C;
  GemmEpilogue ep;
  long long* trace = nullptr;   // optional device buffer [4 roles][64 tiles][4]: clock64 timeline of CTA 0 (mmx_gemm_trace; profiling aid)
};

__device__ __forceinline__ void tmem_st8v(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st16v(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

// (x0, x1) -> packed fp16 pair hi (x0 in the low half: element k of a K-major operand sits below element k+1) and the
// packed pair of the scaled residuals
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((x0 - hf.x) * LO_SCALE, (x1 - hf.y) * LO_SCALE);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int BN, int G, bool ACT>
__global__ void __launch_bounds__(threads_for(G), 1) gemm_f16x3_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                const __grid_constant__ CUtensorMap mapBhi,
                                                                const __grid_constant__ CUtensorMap mapBlo, Params p) {
  using cfg = Cfg<BN>;
  constexpr int SPLIT_WARPS = 4 * G;
  constexpr int B_BYTES = cfg::B_BYTES, B_STAGE = cfg::B_STAGE, SB = cfg::SB, TS = cfg::TS, CW = cfg::CW;
  constexpr uint32_t TM_MAIN = cfg::TM_MAIN, TM_CROSS = cfg::TM_CROSS, TM_A = cfg::TM_A;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // swizzle-128B tiles need 1024 B alignment
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t b_base = smem_base + SA * A_BYTES;                      // W ring behind the A ring
  const uint32_t bar_base = b_base + SB * B_STAGE;
  // barriers (8 B each): full_a[SA], empty_a[SA], full_b[SB], empty_b[SB], split_done[TS], tmem_free[TS], tmem_full, tmem_empty
  auto full_a = [&](int s) { return bar_base + 8u * s; };
  auto empty_a = [&](int s) { return bar_base + 8u * (SA + s); };
  auto full_b = [&](int s) { return bar_base + 8u * (2 * SA + s); };
  auto empty_b = [&](int s) { return bar_base + 8u * (2 * SA + SB + s); };
  auto split_bar = [&](int s) { return bar_base + 8u * (2 * SA + 2 * SB + s); };
  auto tfree_bar = [&](int s) { return bar_base + 8u * (2 * SA + 2 * SB + TS + s); };
  constexpr int NBAR = 2 * SA + 2 * SB + 2 * TS;
  const uint32_t tfull_bar = bar_base + 8u * NBAR, tempty_bar = bar_base + 8u * (NBAR + 1);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + SA * A_BYTES + SB * B_STAGE + 8 * (NBAR + 2));
  float* stg_base = reinterpret_cast<float*>(smem_gen + SA * A_BYTES + SB * B_STAGE + 512);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(full_a(s), 1); mbar_init(empty_a(s), SPLIT_WARPS); }
    for (int s = 0; s < SB; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), 1); }
    for (int s = 0; s < TS; ++s) { mbar_init(split_bar(s), SPLIT_WARPS); mbar_init(tfree_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBlo) : "memory");
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
  // Programmatic dependent launch: everything above (barriers, tensor-map prefetch, TMEM allocation) touched no global
  // data and may run while the kernel before this one in the stream is still draining; its results (A, the residual, the
  // device-side row count) are read only below.  A no-op when the launch carries no programmatic dependency.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // (No early griddepcontrol.launch_dependents: measured 8300 -> 7880 maps/s - the dependent's CTAs park on the SMs the
  // last wave leaves idle, which the OTHER tower's stream would have used.)
  if (p.ep.m_dev) p.M = min(p.M, __ldg(p.ep.m_dev));       // ragged batch: the host-side M is only an upper bound
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int nk = (p.K + BK - 1) / BK;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer, A (whole warp loops, one lane issues)
    int stage = 0; uint32_t phase = 0;
    int itp = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++itp) {
      const int m0 = (t % tiles_m) * BM;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(empty_a(stage), phase ^ 1);
        const uint32_t sa = smem_base + stage * A_BYTES;
        if (elect_one()) {
          if (kb == 0) HX_TRACE(0, itp, 0);
          if (kb == nk - 1) HX_TRACE(0, itp, 1);
          mbar_expect_tx(full_a(stage), A_BYTES);
          tma_load_2d(sa, &mapA, full_a(stage), kb * BK, m0);
          tma_load_2d(sa + A_SUB, &mapA, full_a(stage), kb * BK + 32, m0);
        }
        __syncwarp();
        if (++stage == SA) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ TMA producer, W planes
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int n0 = (t / tiles_m) * BN;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(empty_b(stage), phase ^ 1);
        const uint32_t sb = b_base + stage * B_STAGE;
        if (elect_one()) {
          mbar_expect_tx(full_b(stage), 2 * B_BYTES);
          tma_load_2d(sb, &mapBhi, full_b(stage), kb * BK, n0);
          tma_load_2d(sb + B_BYTES, &mapBlo, full_b(stage), kb * BK, n0);
        }
        __syncwarp();
        if (++stage == SB) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp loops, one lane issues)
    // instruction descriptor: D = f32 (bit 4), A = B = f16 (format 0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    int sb = 0; uint32_t pb = 0;
    int ts = 0; uint32_t pt = 0;
    int it = 0;
    const uint32_t d_main = tmem_base + TM_MAIN, d_cross = tmem_base + TM_CROSS;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      mbar_wait(tempty_bar, (uint32_t)(it & 1) ^ 1);               // epilogue drained the accumulators
      tc_fence_after();
      if (lane == 0) HX_TRACE(1, it, 0);
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(split_bar(ts), pt);                             // A hi/lo of this slab are in TMEM
        mbar_wait(full_b(sb), pb);                                // W planes of this slab landed
        if (lane == 0 && kb == 0) HX_TRACE(1, it, 1);
        if (lane == 0 && kb == nk - 1) HX_TRACE(1, it, 2);
        tc_fence_after();
        const uint32_t sbm = b_base + sb * B_STAGE;
        const uint64_t b_hi = make_desc(sbm), b_lo = make_desc(sbm + B_BYTES);
        const uint32_t a_hi = tmem_base + TM_A + 64u * ts, a_lo = a_hi + 32u;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {                     // UMMA_K = 16 (f16): +32 B in smem (+2 in the descriptor), +8 TMEM columns
            const uint64_t adv = (uint64_t)(2 * k);
            const uint32_t first = (kb > 0 || k > 0) ? 1u : 0u;
            umma_f16_ts(d_cross, a_lo + 8u * k, b_hi + adv, idesc, first);
            umma_f16_ts(d_cross, a_hi + 8u * k, b_lo + adv, idesc, 1u);
            umma_f16_ts(d_main, a_hi + 8u * k, b_hi + adv, idesc, first);
          }
          umma_commit(empty_b(sb));                               // W stage and TMEM A slab are free when these MMAs retire
          umma_commit(tfree_bar(ts));
          if (kb == nk - 1) umma_commit(tfull_bar);               // accumulators complete -> epilogue
        }
        __syncwarp();
        if (++sb == SB) { sb = 0; pb ^= 1; }
        if (++ts == TS) { ts = 0; pt ^= 1; }
      }
    }
  } else if (warp >= SPLIT_WARP0) {
    // ------------------------------------------------------------------ splitter: A slab (smem, raw fp32) -> hi / lo' fp16 -> TMEM
    const int q = warp & 3;                                     // TMEM lane quarter this warp may write
    const int grp = (warp - SPLIT_WARP0) >> 2;                  // which 64 / G K-columns of the slab
    const int row = q * 32 + lane;                              // tile row handled by this thread
    const uint32_t row_off = (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
    constexpr int CH = 16 / G;                                  // 16-byte chunks (4 fp32 -> 2 TMEM columns per plane) per warp and slab
    constexpr int RC = CH < 8 ? CH : 8;                         // chunks per round (one tcgen05.st per plane and round)
    int sa = 0; uint32_t pa = 0;
    int ts = 0; uint32_t pt = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(full_a(sa), pa);
        mbar_wait(tfree_bar(ts), pt ^ 1);                       // the MMAs that read this TMEM slab last time have retired
        tc_fence_after();
        const uint8_t* a_raw = smem_gen + sa * A_BYTES + row_off;
        const uint32_t t_hi = tmem_base + ((uint32_t)(q * 32) << 16) + TM_A + 64u * ts;
#pragma unroll
        for (int r = 0; r < CH / RC; ++r) {
          const int cg0 = grp * CH + r * RC;                    // first chunk of the round: box cg0 >> 3 (k = 32 box ..), chunk cg0 & 7 in it
          uint32_t hi[2 * RC], lo[2 * RC];
#pragma unroll
          for (int c = 0; c < RC; ++c) {                          // chunk c of the box row sits at swizzled position c ^ (row & 7)
            const uint32_t cb = (uint32_t)((cg0 & 7) + c);
            const float4 x = *reinterpret_cast<const float4*>(a_raw + (cg0 >> 3) * A_SUB + ((cb ^ sw) << 4));
            split2(x.x, x.y, hi[2 * c], lo[2 * c]);
            split2(x.z, x.w, hi[2 * c + 1], lo[2 * c + 1]);
          }
          if (r == CH / RC - 1) {                                 // this warp's share of the raw slab is in registers: hand the stage back
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_a(sa));
          }
          if constexpr (RC == 8) {
            tmem_st16v(t_hi + 2u * cg0, hi);
            tmem_st16v(t_hi + 32u + 2u * cg0, lo);
          } else {
            tmem_st8v(t_hi + 2u * cg0, hi);
            tmem_st8v(t_hi + 32u + 2u * cg0, lo);
          }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(split_bar(ts));
        if (++sa == SA) { sa = 0; pa ^= 1; }
        if (++ts == TS) { ts = 0; pt ^= 1; }
      }
    }
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + EPI_WARPS) {
    // ------------------------------------------------------------------ epilogue (8 warps: lane quarter x column half)
    const int q = warp & 3;                                     // TMEM lane quarter this warp may read
    const int ch = (warp - EPI_WARP0) >> 2;                     // column half: CW of the BN accumulator columns
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int m0 = (t % tiles_m) * BM, n0 = (t / tiles_m) * BN;
      mbar_wait(tfull_bar, (uint32_t)(it & 1));
      if (warp == EPI_WARP0 && lane == 0) HX_TRACE(2, it, 0);
      tc_fence_after();
      // drain main + cross accumulators into registers (one RN fma each), then hand TMEM back before any global traffic
      uint32_t acc[CW];
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * CW);
#pragma unroll
      for (int c = 0; c + 16 <= CW; c += 16) {
        uint32_t x[16];
        tmem_ld16_nowait(trow + TM_MAIN + c, acc + c);
        tmem_ld16_nowait(trow + TM_CROSS + c, x);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[c + j] = __float_as_uint(fmaf(__uint_as_float(x[j]), LO_INV, __uint_as_float(acc[c + j])));
      }
      if constexpr (CW % 16 == 8) {
        constexpr int c = CW - 8;
        uint32_t x[8];
        tmem_ld8_nowait(trow + TM_MAIN + c, acc + c);
        tmem_ld8_nowait(trow + TM_CROSS + c, x);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[c + j] = __float_as_uint(fmaf(__uint_as_float(x[j]), LO_INV, __uint_as_float(acc[c + j])));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);                  // next tile's MMAs may start
      if (warp == EPI_WARP0 && lane == 0) HX_TRACE(2, it, 1);
      epilogue_store_rows<CW, ACT>(acc, stg_base + (warp - EPI_WARP0) * EPI_STAGE_FLOATS, m0 + q * 32, n0 + ch * CW, p.M, p.N, p.C, p.ldc, 0, 0,
                                   p.ep, lane);
      if (warp == EPI_WARP0 && lane == 0) HX_TRACE(2, it, 2);
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

// ------------------------------------------------------------------------------------------------------------------
// Backend 3 (selectable, mmx_set_gemm_backend(3)): A arrives as fp16 hi / lo' planes as well - split ONCE per GEMM by
// split_a_planes_kernel instead of once per column tile inside the main loop - and the product runs on CTA PAIRS.
// Measured (profiles/gemm_bench_r2_epilogue.log, CUDA-graph timing): the pair main loop runs at the tensor pipe's floor
// (690 clk per 256 x 128 x 64 slab, profiles/gemm_trace_r2_after.log) and the GEMM alone beats the default kernel by 0-25 %
// (QKV 32.3 vs 31.4 us, fc1 36.1 vs 39.7, text fc2 29.2 vs 43.2), but the separate split pass costs 6-14 us per GEMM, so
// end to end it loses until the producers of A (layer norm, attention output, the GELU epilogue) write the planes
// themselves.  That is the next step for this kernel; until then the in-kernel split above stays the default.
namespace pre {
constexpr int AP_BYTES = BM * BK * 2;                    // one A plane slab (128 x 64 fp16): 16 KB

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair form (cta_group::2): one 256 x BN output tile per PAIR of SMs.  Each CTA loads its own 128 A rows and only HALF
// of the W rows (the MMA reads both halves through the cluster): a 256 x 256 tile costs 64 KB of L2 -> SM traffic per CTA
// and K-slab for 1.6 x the MACs of the 128 x 160 tile's 72 KB - 1.8 x fewer bytes per MAC.  (It was built on the hypothesis
// that the single-CTA kernel was bound by L2 delivery; the timeline showed the epilogue instead - gemm_epilogue.cuh - but
// the pair main loop does run at the tensor pipe's floor.)  TMEM holds up to 256 main + 256 cross columns, which is why A
// comes from shared memory (pre-split planes), not from TMEM; at BN = 128 there are two accumulator sets and the drain
// of one tile overlaps the main loop of the next.
//   both CTAs   warp 0: TMA producer (own A planes, own half of the W planes; completion bytes land on the LEADER's
//               full barrier);  warps 4 ..: epilogue of the CTA's own 128 rows (arrive on the leader's tmem_empty)
//   leader      warp 1: MMA issuer (M = 256 across the pair); tcgen05.commit multicast -> empty / tmem_full of both CTAs
namespace pair2 {
template <int BN> struct CfgQ {
  static_assert(BN % 32 == 0 && BN >= 128 && BN <= 256, "UMMA N for M = 256: multiple of 16; each CTA holds BN / 2 rows of W");
  static constexpr int BNH = BN / 2;
  static constexpr int B_BYTES = BNH * BK * 2;                   // one W plane slab of this CTA
  static constexpr int STAGE = 2 * AP_BYTES + 2 * B_BYTES;
  static constexpr uint32_t CSTRIDE = BN;
  static constexpr int ACC = 4 * BN <= 512 ? 2 : 1;
  static constexpr int EG = BN > 160 ? 4 : 2;                    // epilogue column groups (x 4 lane quarters = epilogue warps)
  static constexpr int CW = BN / EG;
  static constexpr int EPI = 4 * EG;
  static constexpr int EPI_STG = EPI * EPI_STAGE_FLOATS * 4;
  static constexpr int ST = (227 * 1024 - 1024 - 512 - EPI_STG) / STAGE > 6 ? 6 : (227 * 1024 - 1024 - 512 - EPI_STG) / STAGE;
  static constexpr int SMEM_BYTES = ST * STAGE + 1024 + 512 + EPI_STG;
  static constexpr int THREADS = (4 + EPI) * 32;
  static_assert(ST >= 3 && CW % 16 == 0, "ring depth / epilogue chunking");
};

template <int BN, bool ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CfgQ<BN>::THREADS, 1)
gemm_f16x3_pair_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                       const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, Params p) {
  using cfg = CfgQ<BN>;
  constexpr int B_BYTES = cfg::B_BYTES, STAGE = cfg::STAGE, ST = cfg::ST, CW = cfg::CW, ACC = cfg::ACC, EPI = cfg::EPI, BNH = cfg::BNH;
  constexpr uint32_t CSTRIDE = cfg::CSTRIDE;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + ST * STAGE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };               // used on the leader only
  auto empty_bar = [&](int s) { return bar_base + 8u * (ST + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * ST + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * ST + ACC + a); };   // used on the leader only
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + ST * STAGE + 8 * (2 * ST + 2 * ACC));
  float* stg_base = reinterpret_cast<float*>(smem_gen + ST * STAGE + 512);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.ep.m_dev) p.M = min(p.M, __ldg(p.ep.m_dev));
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int tiles_m = (p.M + 2 * BM - 1) / (2 * BM), tiles_n = (p.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int nk = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < ACC; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * EPI); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAlo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBlo) : "memory");
  }
  cluster_sync();                                                // both CTAs' barriers exist before any remote arrive
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    int stage = 0; uint32_t phase = 0;
    int itp = 0;
    for (int t = pair; t < num_tiles; t += num_pairs, ++itp) {
      const int m0 = (t % tiles_m) * 2 * BM + (int)rank * BM, n0 = (t / tiles_m) * BN + (int)rank * BNH;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sa = smem_base + stage * STAGE;
        if (elect_one()) {
          if (kb == 0) HX_TRACE(0, itp, 0);
          if (kb == nk - 1) HX_TRACE(0, itp, 1);
          const uint32_t fb = mapa(full_bar(stage), 0);            // the leader's barrier collects the bytes of both CTAs
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * STAGE);
          tma_load_2d_pair(sa, &mapAhi, fb, kb * BK, m0);
          tma_load_2d_pair(sa + AP_BYTES, &mapAlo, fb, kb * BK, m0);
          tma_load_2d_pair(sa + 2 * AP_BYTES, &mapBhi, fb, kb * BK, n0);
          tma_load_2d_pair(sa + 2 * AP_BYTES + B_BYTES, &mapBlo, fb, kb * BK, n0);
        }
        __syncwarp();
        if (++stage == ST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (rank == 0) {
      constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);   // M = 256 across the pair
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
        const int as = it % ACC;
        mbar_wait(tempty_bar(as), (uint32_t)((it / ACC) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)as * 2u * CSTRIDE, d_cross = d_main + CSTRIDE;
        if (lane == 0) HX_TRACE(1, it, 0);
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(full_bar(stage), phase);
          if (lane == 0 && kb == 0) HX_TRACE(1, it, 1);
          if (lane == 0 && kb == nk - 1) HX_TRACE(1, it, 2);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE;
          const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + AP_BYTES);
          const uint64_t b_hi = make_desc(sa + 2 * AP_BYTES), b_lo = make_desc(sa + 2 * AP_BYTES + B_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t adv = (uint64_t)(2 * k);
              const uint32_t first = (kb > 0 || k > 0) ? 1u : 0u;
              umma_f16_ss_pair(d_cross, a_lo + adv, b_hi + adv, idesc, first);
              umma_f16_ss_pair(d_cross, a_hi + adv, b_lo + adv, idesc, 1u);
              umma_f16_ss_pair(d_main, a_hi + adv, b_hi + adv, idesc, first);
            }
            umma_commit_pair(empty_bar(stage));                   // both CTAs' stage is free when these MMAs retire
            if (kb == nk - 1) umma_commit_pair(tfull_bar(as));
          }
          __syncwarp();
          if (++stage == ST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + EPI) {
    // ------------------------------------------------------------------ epilogue: this CTA's 128 rows (lane quarter x column group)
    const int q = warp & 3;
    const int ch = (warp - EPI_WARP0) >> 2;
    int it = 0;
    for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
      const int m0 = (t % tiles_m) * 2 * BM + (int)rank * BM, n0 = (t / tiles_m) * BN;
      const int as = it % ACC;
      mbar_wait(tfull_bar(as), (uint32_t)((it / ACC) & 1));
      if (warp == EPI_WARP0 && lane == 0) HX_TRACE(2, it, 0);
      tc_fence_after();
      uint32_t acc[CW];
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * 2u * CSTRIDE + (uint32_t)(ch * CW);
#pragma unroll
      for (int c = 0; c < CW; c += 16) {
        uint32_t x[16];
        tmem_ld16_nowait(trow + c, acc + c);
        tmem_ld16_nowait(trow + CSTRIDE + c, x);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[c + j] = __float_as_uint(fmaf(__uint_as_float(x[j]), LO_INV, __uint_as_float(acc[c + j])));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa(tempty_bar(as), 0));
      if (warp == EPI_WARP0 && lane == 0) HX_TRACE(2, it, 1);
      epilogue_store_rows<CW, ACT>(acc, stg_base + (warp - EPI_WARP0) * EPI_STAGE_FLOATS, m0 + q * 32, n0 + ch * CW, p.M, p.N, p.C, p.ldc, 0, 0,
                                   p.ep, lane);
      if (warp == EPI_WARP0 && lane == 0) HX_TRACE(2, it, 2);
    }
  }
  tc_fence_before();
  cluster_sync();                                                // the peer may read this CTA's smem / signal its barriers until here
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}
}  // namespace pair2

// fp32 A [M, K] (row stride lda, lda % 4 == 0) -> hi / lo' fp16 planes [M, ldp]; 4 elements per thread.  Columns K .. ldp-1
// are never read (the tensor maps declare K columns, TMA zero-fills beyond).
__global__ void __launch_bounds__(256) split_a_planes_kernel(const float* __restrict__ A, int lda, __half* __restrict__ hi,
                                                             __half* __restrict__ lo, int ldp, int M, int K) {
  const int kc = (K + 3) / 4;
  const long long total = (long long)M * kc;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / kc), k = (int)(i % kc) * 4;
    const float* src = A + (long long)m * lda + k;
    float4 x;
    if (k + 4 <= K) x = *reinterpret_cast<const float4*>(src);
    else { x.x = src[0]; x.y = k + 1 < K ? src[1] : 0.f; x.z = k + 2 < K ? src[2] : 0.f; x.w = 0.f; }
    uint2 h, l;
    split2(x.x, x.y, h.x, l.x);
    split2(x.z, x.w, h.y, l.y);
    *reinterpret_cast<uint2*>(hi + (long long)m * ldp + k) = h;       // ldp % 8 == 0 and k % 4 == 0: 8-byte aligned, inside the row
    *reinterpret_cast<uint2*>(lo + (long long)m * ldp + k) = l;
  }
}
}  // namespace pre

// fp32 [N,K] (row stride ldw) -> hi / lo' fp16 planes [N, ldp], pad columns K .. ldp-1 zero
__global__ void __launch_bounds__(256) pack_f16x3_kernel(const float* __restrict__ W, int ldw, __half* __restrict__ hi,
                                                         __half* __restrict__ lo, int ldp, int N, int K) {
  const long long total = (long long)N * ldp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / ldp), k = (int)(i % ldp);
    float x = 0.f;
    if (k < K) x = W[(long long)n * ldw + k];
    const __half h = __float2half_rn(x);
    hi[i] = h;
    lo[i] = __float2half_rn((x - __half2float(h)) * LO_SCALE);
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static long long* g_trace = nullptr;

static int make_map(CUtensorMap* map, CUtensorMapDataType dt, int esize, const void* base, int rows, int K, int ld, int box_k,
                    int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esize};
  cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return 1;
  }
  return 0;
}

static int ensure_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr) {
    set_error("cuTensorMapEncodeTiled not available");
    return 1;
  }
  g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  return 0;
}

// Tile width for this launch: the BN whose tile count costs the fewest SM-waves x BN (ties go to the narrower tile).
// A ragged launch carries the host's upper bound of the row count; the choice never changes a result bit.
static int pick_bn(int M, int N) {
  const int forced = gemm_tc_tile_n(-1);
  if (forced) return forced;
  const int sms = sm_count(), tiles_m = cdiv(M, BM);
  int best = 128;
  long long best_cost = 1ll << 60;
  for (int bn : {128, 144, 160}) {
    const long long cost = (long long)cdiv(tiles_m * cdiv(N, bn), sms) * bn;
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

template <int BN, int G>
static int launch_bn_g(const float* A, int lda, const BOperand& B, int M, int N, int K, const Params& p, cudaStream_t st) {
  using cfg = Cfg<BN>;
  CUtensorMap mapA, mapBhi, mapBlo;
  MMX_TRY(make_map(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, A, M, K, lda, 32, BM));
  MMX_TRY(make_map(&mapBhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, B.hi, N, K, B.ldp, BK, BN));
  MMX_TRY(make_map(&mapBlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, B.lo, N, K, B.ldp, BK, BN));
  // per device: the attribute belongs to the (function, device) pair
  const int tiles = cdiv(M, BM) * cdiv(N, BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  // launched as a programmatic dependent of whatever precedes it in the stream: the prologue overlaps that kernel's tail
  static int pdl = -1;
  if (pdl < 0) { const char* e = getenv("MMX_PDL"); pdl = e ? atoi(e) != 0 : 1; }
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(grid); lc.blockDim = dim3(threads_for(G)); lc.dynamicSmemBytes = cfg::SMEM_BYTES; lc.stream = st;
  lc.attrs = attr; lc.numAttrs = pdl ? 1 : 0;
  if (p.ep.pre || p.ep.C_act) {
    MMX_CHECK_CUDA(cudaFuncSetAttribute(gemm_f16x3_kernel<BN, G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::SMEM_BYTES));
    MMX_CHECK_CUDA(cudaLaunchKernelEx(&lc, gemm_f16x3_kernel<BN, G, true>, mapA, mapBhi, mapBlo, p));
  } else {
    MMX_CHECK_CUDA(cudaFuncSetAttribute(gemm_f16x3_kernel<BN, G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::SMEM_BYTES));
    MMX_CHECK_CUDA(cudaLaunchKernelEx(&lc, gemm_f16x3_kernel<BN, G, false>, mapA, mapBhi, mapBlo, p));
  }
  MMX_LAUNCH_CHECK();
  return 0;
}

template <int BN>
static int launch_bn(const float* A, int lda, const BOperand& B, int M, int N, int K, const Params& p, cudaStream_t st) {
  return launch_bn_g<BN, 1>(A, lda, B, M, N, K, p, st);   // G = 2 / 4 were measured: no gain (the split is not the limiter)
}

template <int BN>
static int launch_pair(const __half* Ahi, const __half* Alo, int ldpa, const BOperand& B, int M, int N, int K, const Params& p,
                       cudaStream_t st) {
  using cfg = pre::pair2::CfgQ<BN>;
  CUtensorMap mapAhi, mapAlo, mapBhi, mapBlo;
  MMX_TRY(make_map(&mapAhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, Ahi, M, K, ldpa, BK, BM));
  MMX_TRY(make_map(&mapAlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, Alo, M, K, ldpa, BK, BM));
  MMX_TRY(make_map(&mapBhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, B.hi, N, K, B.ldp, BK, cfg::BNH));
  MMX_TRY(make_map(&mapBlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, B.lo, N, K, B.ldp, BK, cfg::BNH));
  const int tiles = cdiv(M, 2 * BM) * cdiv(N, BN);
  const int max_pairs = sm_count() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  if (p.ep.pre || p.ep.C_act) {      // __cluster_dims__(2,1,1)
    MMX_CHECK_CUDA(cudaFuncSetAttribute(pre::pair2::gemm_f16x3_pair_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::SMEM_BYTES));
    pre::pair2::gemm_f16x3_pair_kernel<BN, true><<<2 * pairs, cfg::THREADS, cfg::SMEM_BYTES, st>>>(mapAhi, mapAlo, mapBhi, mapBlo, p);
  } else {
    MMX_CHECK_CUDA(cudaFuncSetAttribute(pre::pair2::gemm_f16x3_pair_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::SMEM_BYTES));
    pre::pair2::gemm_f16x3_pair_kernel<BN, false><<<2 * pairs, cfg::THREADS, cfg::SMEM_BYTES, st>>>(mapAhi, mapAlo, mapBhi, mapBlo, p);
  }
  MMX_LAUNCH_CHECK();
  return 0;
}

// Tile width of the pair kernel: estimated clocks = waves x (K-slabs x slab time + exposed accumulator hand-over).  Slab
// time and hand-over from the clock64 timeline (profiles/gemm_trace_r2_after.log): 690 clk per slab at BN = 128 (two
// accumulator sets: the drain is hidden), 1340 at BN = 256 plus ~5500 clk between the last MMA and the next tile's first.
static int pick_bn_pair(int M, int N, int K) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("MMX_PAIR_BN"); forced = e ? atoi(e) : 0; }
  if (forced == 128 || forced == 160 || forced == 192 || forced == 256) return forced;
  const int pairs = sm_count() / 2, tiles_m = cdiv(M, 2 * BM), nk = cdiv(K, BK);
  int best = 128;
  long long best_cost = 1ll << 60;
  for (int bn : {128, 160, 192, 256}) {
    const long long tiles = (long long)tiles_m * cdiv(N, bn);
    const long long waves = (tiles + pairs - 1) / pairs;
    const long long slab = 690ll * bn / 128;
    const long long handover = 4 * bn <= 512 ? 800 : 1500 + 16ll * bn;
    const long long cost = waves * (nk * slab + handover);
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

// split A once (stream-ordered scratch, same bytes as the fp32 original), then the TMA -> MMA pair kernel
static int gemm_presplit_pair(const float* A, int lda, const BOperand& B, int M, int N, int K, const Params& p, cudaStream_t st) {
  const int ldpa = round_up(K, 8);
  __half* planes = nullptr;
  keep_stream_scratch_cached();
  MMX_CHECK_CUDA(cudaMallocAsync((void**)&planes, (size_t)2 * M * ldpa * sizeof(__half), st));
  __half *Ahi = planes, *Alo = planes + (size_t)M * ldpa;
  const long long total = (long long)M * ((K + 3) / 4);
  const int grid = (int)((total + 255) / 256 < (long long)sm_count() * 8 ? (total + 255) / 256 : (long long)sm_count() * 8);
  pre::split_a_planes_kernel<<<grid, 256, 0, st>>>(A, lda, Ahi, Alo, ldpa, M, K);
  MMX_LAUNCH_CHECK();
  int rc = 0;
  switch (pick_bn_pair(M, N, K)) {
    case 160: rc = launch_pair<160>(Ahi, Alo, ldpa, B, M, N, K, p, st); break;
    case 192: rc = launch_pair<192>(Ahi, Alo, ldpa, B, M, N, K, p, st); break;
    case 256: rc = launch_pair<256>(Ahi, Alo, ldpa, B, M, N, K, p, st); break;
    default: rc = launch_pair<128>(Ahi, Alo, ldpa, B, M, N, K, p, st); break;
  }
  cudaFreeAsync(planes, st);
  return rc;
}

}  // namespace hx

void gemm_f16x3_set_trace(long long* buf) { hx::g_trace = buf; }

int pack_f16x3_ld(int K) { return round_up(K, 8); }
size_t pack_f16x3_bytes(int N, int K) { return (size_t)2 * N * pack_f16x3_ld(K) * sizeof(uint16_t); }

int pack_f16x3(const float* W, int ldw, int N, int K, void* packed, cudaStream_t st) {
  if (N == 0 || K == 0) return 0;
  MMX_REQUIRE(aligned16(packed), "packed operand must be 16-byte aligned");
  const int ldp = pack_f16x3_ld(K);
  __half* hi = reinterpret_cast<__half*>(packed);
  __half* lo = hi + (size_t)N * ldp;
  const long long total = (long long)N * ldp;
  const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  hx::pack_f16x3_kernel<<<grid, 256, 0, st>>>(W, ldw, hi, lo, ldp, N, K);
  MMX_LAUNCH_CHECK();
  return 0;
}

BOperand packed_operand(const float* W, int ldw, const void* packed, int N, int K) {
  BOperand B;
  B.w = W; B.ldw = ldw;
  if (packed) {
    B.ldp = pack_f16x3_ld(K);
    B.hi = reinterpret_cast<const uint16_t*>(packed);
    B.lo = B.hi + (size_t)N * B.ldp;
  }
  return B;
}

// Which problems go to this kernel.  Independent of M (a sample alone and inside a batch take the same arithmetic path).
bool gemm_f16x3_shape_ok(const float* A, int lda, const BOperand& B, float* C, int ldc, int N, int K, const GemmEpilogue& ep) {
  if (!gemm_tc_available() || !B.hi || !B.lo) return false;
  if (N < 128 || K < 64 || (N % 4) || (lda % 4) || (B.ldp % 8) || (ldc % 4)) return false;   // K itself is free: TMA zero-fills the tail
  if (!aligned16(A) || !aligned16(B.hi) || !aligned16(B.lo) || !aligned16(C)) return false;
  if ((ep.bias && !aligned16(ep.bias)) || (ep.pre && (!aligned16(ep.pre) || ep.ldpre % 4)) ||
      (ep.residual && (!aligned16(ep.residual) || ep.ldres % 4)) || (ep.C_act && !aligned16(ep.C_act)))
    return false;
  return true;
}

int gemm_nt_f16x3(const float* A, int lda, const BOperand& B, float* C, int ldc, int M, int N, int K, const GemmEpilogue& ep,
                  cudaStream_t st, bool pair) {
  if (M == 0 || N == 0) return 0;
  MMX_TRY(hx::ensure_encode());
  hx::Params p{M, N, K, ldc, C, ep};
  p.trace = hx::g_trace;
  if (pair) return hx::gemm_presplit_pair(A, lda, B, M, N, K, p, st);
  switch (hx::pick_bn(M, N)) {
    case 144: return hx::launch_bn<144>(A, lda, B, M, N, K, p, st);
    case 160: return hx::launch_bn<160>(A, lda, B, M, N, K, p, st);
    default: return hx::launch_bn<128>(A, lda, B, M, N, K, p, st);
  }
}

}  // namespace mmx
