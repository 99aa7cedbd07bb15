// Attention forward that stages A = softmax(QK^T) to HBM, and the attention backward that stages dA = dO V^T
// (the two tensors the reference captures with forward / backward hooks) and continues to dQ, dK, dV.
// fp16x3 mma.sync (m16n8k16) products with fp32 softmax; one CTA = one (batch, head, 64-query tile), 8 warps; score rows live in shared memory so any S <= ~1500
// (DETR 850, ViT-L/14@336 577) is handled without a second pass.  Deterministic (no atomics): dK/dV come from a
// second kernel that walks the query tiles for one key tile.
#include "mmx_common.cuh"
#include <math_constants.h>
#include <cuda_fp16.h>

namespace mmx {

// Ragged self-attention (packed rows): sample b owns rows offs[b] .. offs[b]+lens[b]-1 of Q/K/V/O and attends within
// them; the staged A / dA planes stay dense [B,H,T,ld] with rows and columns >= lens[b] zero-filled.
struct Ragged {
  const int* offs = nullptr;
  const int* lens = nullptr;
};

// Query rows per CTA: TQ = 64 (four m16 row blocks, 8 warps: one CTA covers a whole CLIP ViT-B/32 head of 50 rows)
// while the score rows fit in shared memory (S <= ~730), TQ = 32 (4 warps) beyond that (DETR at 800x1066: 850 keys).
// A CTA has TQ/8 warps; warp w owns row block w % (TQ/16) and half w / (TQ/16) of the keys (scores) or of d (P.V).
constexpr int TKEY = 64;  // keys per shared-memory tile
constexpr int KV_THREADS = 256;

// The small matrix products per head (scores, P.V and their backward twins) run on the tensor cores as
// mma.sync.m16n8k16 with fp16 operands and the same fp32-faithful 3-pass split as the linear GEMMs (gemm_f16x3.cu):
// x = hi + lo' / 2048 with hi = fp16(x), lo' = fp16((x - hi) * 2048); hi*hi goes into one fp32 accumulator, lo'*hi + hi*lo'
// into a second one that is folded in once (x 1/2048) at the end.  fp16 x fp16 products are exact in the fp32 accumulator,
// one instruction covers K = 16 (m16n8k8 TF32: 8) at twice the TF32 rate, so against the 3xTF32 version this halves the
// MMA count and the fragment loads per unit of K.  Operand range: |x| <= 65504 and an absolute floor of 3e-11 - activations,
// probabilities and the (power-of-two normalised, see gscale) gradient stream sit well inside.  tcgen05 / TMEM is the wrong
// tool here: a head is 50x50x64 (or 77x77x64), far below one 128-row UMMA tile.
//
// Fragment layout of m16n8k16 (g = lane / 4, t = lane % 4), every register = two consecutive-k fp16 values:
//   A (16x16, row)  a0 (g, 2t..2t+1)  a1 (g+8, 2t..)  a2 (g, 2t+8..)  a3 (g+8, 2t+8..)
//   B (16x8, col)   b0 (k = 2t..2t+1, n = g)  b1 (k = 2t+8.., n = g)
//   C (16x8)        c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1)
// Shared memory keeps fp32; the split happens at fragment-load time.  Strides are chosen so that each fragment load is
// bank-conflict free: operands indexed [row or n = g][k = 2t] are read with 64-bit loads and want a stride = 8 (mod 16)
// floats; operands indexed [k = 2t][n or row = g] are read with 32-bit loads from rows 2t and 2t+1 and want a stride
// = 4 (mod 16).
constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;
__device__ __forceinline__ void split_f16(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);            // x0 in the low half: element k sits below element k+1
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((x0 - hf.x) * LO_SCALE, (x1 - hf.y) * LO_SCALE);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
struct FragA { uint32_t hi[4], lo[4]; };
struct FragB { uint32_t hi[2], lo[2]; };
// A fragment of a [row][k] operand (stride ld): p points at (row0 + g, k0 + 2t); `mul` scales the operand before the split
__device__ __forceinline__ FragA frag_a_rowmajor(const float* p, int ld, float mul = 1.f) {
  FragA f;
  const float2 x0 = *reinterpret_cast<const float2*>(p), x1 = *reinterpret_cast<const float2*>(p + 8 * ld);
  const float2 x2 = *reinterpret_cast<const float2*>(p + 8), x3 = *reinterpret_cast<const float2*>(p + 8 * ld + 8);
  split_f16(x0.x * mul, x0.y * mul, f.hi[0], f.lo[0]);
  split_f16(x1.x * mul, x1.y * mul, f.hi[1], f.lo[1]);
  split_f16(x2.x * mul, x2.y * mul, f.hi[2], f.lo[2]);
  split_f16(x3.x * mul, x3.y * mul, f.hi[3], f.lo[3]);
  return f;
}
// A fragment of a TRANSPOSED operand stored [k][row] (stride ld): p points at (k0 + 2t, row0 + g)
__device__ __forceinline__ FragA frag_a_kmajor(const float* p, int ld) {
  FragA f;
  split_f16(p[0], p[ld], f.hi[0], f.lo[0]);
  split_f16(p[8], p[ld + 8], f.hi[1], f.lo[1]);
  split_f16(p[8 * ld], p[9 * ld], f.hi[2], f.lo[2]);
  split_f16(p[8 * ld + 8], p[9 * ld + 8], f.hi[3], f.lo[3]);
  return f;
}
// B fragment of an operand stored [n][k] (stride ld): p points at (n0 + g, k0 + 2t)
__device__ __forceinline__ FragB frag_b_nmajor(const float* p) {
  FragB f;
  const float2 x0 = *reinterpret_cast<const float2*>(p), x1 = *reinterpret_cast<const float2*>(p + 8);
  split_f16(x0.x, x0.y, f.hi[0], f.lo[0]);
  split_f16(x1.x, x1.y, f.hi[1], f.lo[1]);
  return f;
}
// B fragment of an operand stored [k][n] (stride ld): p points at (k0 + 2t, n0 + g)
__device__ __forceinline__ FragB frag_b_kmajor(const float* p, int ld) {
  FragB f;
  split_f16(p[0], p[ld], f.hi[0], f.lo[0]);
  split_f16(p[8 * ld], p[9 * ld], f.hi[1], f.lo[1]);
  return f;
}
__device__ __forceinline__ void mma3(float (&main)[4], float (&cross)[4], const FragA& a, const FragB& b) {
  mma_f16(cross, a.lo, b.hi);
  mma_f16(cross, a.hi, b.lo);
  mma_f16(main, a.hi, b.hi);
}

__host__ __device__ inline int score_ld(int S) { return round_up(S, 16) + 8; }  // smem stride of a score row: = 8 (mod 16)

template <int HD>
struct AttnSmem {
  static constexpr int LDX = HD + 8;   // [row / key][d] operands (Q, dO tiles; K, V as the scores' B operand): = 8 (mod 16)
  static constexpr int LDV = HD + 4;   // [key][d] operand of the P.V product: = 4 (mod 16)
  static size_t bytes(int S, int TQ) { return sizeof(float) * ((size_t)TQ * score_ld(S) + (size_t)TQ * LDX + (size_t)TKEY * LDX); }
};

// 16-byte asynchronous global -> shared copy (LDGSTS); !valid zero-fills the destination without touching src
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// rows r0 .. r0+ROWS-1 of Y (zero beyond `limit`) -> sY with row stride ld, asynchronously (commit + wait by the caller)
template <int HD, int ROWS>
__device__ __forceinline__ void load_rows_async(const float* __restrict__ Y, int ldy, long long ybase, int r0, int limit,
                                                float* sY, int ld) {
  for (int e = threadIdx.x; e < ROWS * (HD / 4); e += blockDim.x) {
    const int r = e / (HD / 4), d = (e % (HD / 4)) * 4;
    const bool ok = r0 + r < limit;
    cp_async16(sY + r * ld + d, ok ? Y + ybase + (long long)(r0 + r) * ldy + d : Y, ok);
  }
}

// scores[i][j] (i in the 64-row tile, j in [0,S)) = post_scale * sum_d (pre_scale * X[i][d]) * Y[j][d].  The caller has
// issued (cp.async, committed) the X tile into sX and the FIRST key tile into sY; later key tiles stream through sY.
// Warp w owns row block w&3 and the 32-key half w>>2 of each key tile.
template <int HD, int MB>
__device__ __forceinline__ void tile_scores(const float* __restrict__ Y, int ldy, long long ybase, int S,
                                            const float* sX, float* sY, float* sP, int ldP, float pre_scale, float post_scale,
                                            int live_rows) {
  constexpr int LDX = AttnSmem<HD>::LDX;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int m0 = (warp % MB) * 16, kh = (warp / MB) * 32;
  for (int j0 = 0; j0 < S; j0 += TKEY) {
    if (j0 > 0) {
      __syncthreads();
      load_rows_async<HD, TKEY>(Y, ldy, ybase, j0, S, sY, LDX);
      cp_async_commit();
    }
    cp_async_wait_all();
    __syncthreads();
    float acc[4][4] = {}, crs[4][4] = {};
    // live 8-key blocks of this warp's half (may be <= 0); a row block entirely past the sample's rows does no math
    const int ntiles = m0 < live_rows ? (S - j0 - kh + 7) >> 3 : 0;
#pragma unroll 2
    for (int k0 = 0; k0 < HD; k0 += 16) {
      // the reference scales q before the product (auxilary.py:173)
      const FragA a = frag_a_rowmajor(sX + (m0 + g) * LDX + k0 + 2 * t, LDX, pre_scale);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        if (nt < ntiles) {
          const FragB b = frag_b_nmajor(sY + (kh + nt * 8 + g) * LDX + k0 + 2 * t);
          mma3(acc[nt], crs[nt], a, b);
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int j = j0 + kh + nt * 8 + 2 * t;           // even; ldP is even
      if (j < ldP) {
        float* p0 = sP + (m0 + g) * ldP + j;
        *reinterpret_cast<float2*>(p0) = make_float2(fmaf(crs[nt][0], LO_INV, acc[nt][0]) * post_scale,
                                                     fmaf(crs[nt][1], LO_INV, acc[nt][1]) * post_scale);
        *reinterpret_cast<float2*>(p0 + 8 * ldP) = make_float2(fmaf(crs[nt][2], LO_INV, acc[nt][2]) * post_scale,
                                                               fmaf(crs[nt][3], LO_INV, acc[nt][3]) * post_scale);
      }
    }
  }
  __syncthreads();
}

// out[i][d] = sum_j sP[i][j] * Y[j][d].  Warp w owns row block w&3 and the d half w>>2; its HD/16 accumulator
// fragments hold rows m0+g, m0+g+8 and columns d0 + 8*nt + 2t, +1.  The caller has issued the first Y tile into sY
// (stride LDV) with cp.async and passed a __syncthreads after the last write of sP.
template <int HD, int MB>
__device__ __forceinline__ void tile_pv(const float* __restrict__ Y, int ldy, long long ybase, int S, const float* sP,
                                        int ldP, float* sY, float (&out)[HD / 16][4], int live_rows) {
  constexpr int LDV = AttnSmem<HD>::LDV, NT = HD / 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int m0 = (warp % MB) * 16, d0 = (warp / MB) * (HD / 2);
  float crs[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int c = 0; c < 4; ++c) out[nt][c] = crs[nt][c] = 0.f;
  for (int j0 = 0; j0 < S; j0 += TKEY) {
    if (j0 > 0) {
      __syncthreads();
      load_rows_async<HD, TKEY>(Y, ldy, ybase, j0, S, sY, LDV);
      cp_async_commit();
    }
    cp_async_wait_all();
    __syncthreads();
    // keys jn .. round_up(jn, 16) are zero rows of sY, finite columns of sP; dead row blocks skip the math
    const int jn = m0 < live_rows ? min(TKEY, S - j0) : 0;
#pragma unroll 2
    for (int kk = 0; kk < jn; kk += 16) {
      const FragA a = frag_a_rowmajor(sP + (m0 + g) * ldP + j0 + kk + 2 * t, ldP);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const FragB b = frag_b_kmajor(sY + (kk + 2 * t) * LDV + d0 + nt * 8 + g, LDV);
        mma3(out[nt], crs[nt], a, b);
      }
    }
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int c = 0; c < 4; ++c) out[nt][c] = fmaf(crs[nt][c], LO_INV, out[nt][c]);
}

// stores the P.V fragments of tile_pv: rows i0 + m0 + g (+8), columns h*HD + d0 + 8*nt + 2t
template <int HD, int MB>
__device__ __forceinline__ void store_pv(float* __restrict__ O, int ldo, long long row0, int i0, int T, int h,
                                         const float (&out)[HD / 16][4], float mul) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int m0 = (warp % MB) * 16, d0 = (warp / MB) * (HD / 2);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int i = i0 + m0 + g + 8 * half;
    if (i >= T) continue;
    float* orow = O + (row0 + i) * ldo + h * HD + d0 + 2 * t;
#pragma unroll
    for (int nt = 0; nt < HD / 16; ++nt)
      *reinterpret_cast<float2*>(orow + nt * 8) = make_float2(out[nt][2 * half] * mul, out[nt][2 * half + 1] * mul);
  }
}

template <int HD, int TQ>
__global__ void __launch_bounds__(TQ * 4) attention_fwd_kernel(
    const float* __restrict__ Q, int ldq, const float* __restrict__ K, int ldk, const float* __restrict__ V, int ldv,
    const float* __restrict__ key_bias, float* __restrict__ A, int ldA, float* __restrict__ O, int ldo, int H, int T, int S,
    float scale, int flags, Ragged rg) {
  constexpr int LDH = AttnSmem<HD>::LDX, ATT_THREADS = TQ * 4, ATT_WARPS = TQ / 8, MB = TQ / 16;
  extern __shared__ float smem[];
  const int S_pad = score_ld(S);  // shared-memory stride of a score row (from the dense S, also for ragged samples)
  float* sP = smem;
  float* sQ = sP + (size_t)TQ * S_pad;
  float* sKV = sQ + TQ * LDH;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * TQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Tm = T;                                   // height of the staged plane
  long long qrow0 = (long long)b * T, krow0 = (long long)b * S;
  if (rg.lens) { T = S = rg.lens[b]; qrow0 = krow0 = rg.offs[b]; }
  if (i0 >= T) {                                      // tile entirely past this sample's rows: zero its A rows
    for (int e = tid; e < TQ * ldA; e += ATT_THREADS) {
      const int i = i0 + e / ldA;
      if (i < Tm) A[(((long long)b * H + h) * Tm + i) * ldA + e % ldA] = 0.f;
    }
    return;
  }
  const bool scale_scores = flags & MMX_ATTN_SCALE_SCORES;
  load_rows_async<HD, TQ>(Q, ldq, qrow0 * ldq + h * HD, i0, T, sQ, LDH);                 // Q tile and the first K tile together
  load_rows_async<HD, TKEY>(K, ldk, krow0 * ldk + h * HD, 0, S, sKV, LDH);
  cp_async_commit();
  tile_scores<HD, MB>(K, ldk, krow0 * ldk + h * HD, S, sQ, sKV, sP, S_pad, scale_scores ? 1.f : scale, scale_scores ? scale : 1.f,
                  T - i0);
  load_rows_async<HD, TKEY>(V, ldv, krow0 * ldv + h * HD, 0, S, sKV, AttnSmem<HD>::LDV);     // lands during the softmax
  cp_async_commit();
  // softmax per row (warp w owns rows w*8 .. w*8+7), stage A
  for (int rr = 0; rr < TQ / ATT_WARPS; ++rr) {
    const int r = warp * (TQ / ATT_WARPS) + rr, i = i0 + r;
    if (i >= T) {
      if (i < Tm) for (int j = lane; j < ldA; j += 32) A[(((long long)b * H + h) * Tm + i) * ldA + j] = 0.f;
      continue;
    }
    float* row = sP + (size_t)r * S_pad;
    float mx = -CUDART_INF_F;
    for (int j = lane; j < S; j += 32) {
      float v = row[j];
      if (key_bias) v += key_bias[(long long)b * S + j];
      if ((flags & MMX_ATTN_CAUSAL) && j > i) v = -CUDART_INF_F;
      row[j] = v;
      mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    // A key whose bias is -inf is REMOVED (exp(-inf) = 0 exactly, the same bits a -10000 mask gives); a row whose keys are
    // all removed is the softmax over an empty set: A row = 0, output 0, like the reference's empty tensors when the
    // perturbation drivers drop every box (lxmert/lxmert/perturbation.py:110-117 at step 1.0).
    const float mref = mx == -CUDART_INF_F ? 0.f : mx;
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) { const float e = expf(row[j] - mref); row[j] = e; sum += e; }
    sum = warp_sum(sum);
    float* arow = A + (((long long)b * H + h) * Tm + i) * ldA;
    for (int j = lane; j < ldA; j += 32) {
      float p = 0.f;
      if (j < S) { p = sum > 0.f ? row[j] / sum : 0.f; row[j] = p; }
      arow[j] = p;
    }
  }
  __syncthreads();
  float out[HD / 16][4];
  tile_pv<HD, MB>(V, ldv, krow0 * ldv + h * HD, S, sP, S_pad, sKV, out, T - i0);
  store_pv<HD, MB>(O, ldo, qrow0, i0, T, h, out, 1.f);
}

// backward, query side: dA (staged), delta, dQ
template <int HD, int TQ>
__global__ void __launch_bounds__(TQ * 4) attention_bwd_q_kernel(
    const float* __restrict__ dO, int lddo, const float* __restrict__ K, int ldk, const float* __restrict__ V, int ldv,
    const float* __restrict__ A, float* __restrict__ dA, int ldA, float* __restrict__ delta, float* __restrict__ dQ,
    int lddq, int H, int T, int S, float scale, Ragged rg, const float* __restrict__ gscale) {
  constexpr int LDH = AttnSmem<HD>::LDX, ATT_THREADS = TQ * 4, ATT_WARPS = TQ / 8, MB = TQ / 16;
  extern __shared__ float smem[];
  const int S_pad = score_ld(S);
  float* sP = smem;
  float* sX = sP + (size_t)TQ * S_pad;
  float* sKV = sX + TQ * LDH;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * TQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Tm = T;
  long long qrow0 = (long long)b * T, krow0 = (long long)b * S;
  if (rg.lens) { T = S = rg.lens[b]; qrow0 = krow0 = rg.offs[b]; }
  if (i0 >= T) {                                      // past this sample's rows: the hooked gradient is zero there
    for (int e = tid; e < TQ * ldA; e += ATT_THREADS) {
      const int i = i0 + e / ldA;
      if (i < Tm) dA[(((long long)b * H + h) * Tm + i) * ldA + e % ldA] = 0.f;
    }
    return;
  }
  // The gradient stream of sample b may carry a power-of-two factor gscale[b] (fp16x3 GEMM range, see gemm_f16x3.cu);
  // the staged dA is divided by it (exact), everything that continues the backward (delta, dS, dQ) stays scaled.
  const float ginv = gscale ? 1.f / gscale[b] : 1.f;
  load_rows_async<HD, TQ>(dO, lddo, qrow0 * lddo + h * HD, i0, T, sX, LDH);
  load_rows_async<HD, TKEY>(V, ldv, krow0 * ldv + h * HD, 0, S, sKV, LDH);
  cp_async_commit();
  tile_scores<HD, MB>(V, ldv, krow0 * ldv + h * HD, S, sX, sKV, sP, S_pad, 1.f, 1.f, T - i0);
  if (dQ != nullptr) {                                                                    // lands while dA is staged
    load_rows_async<HD, TKEY>(K, ldk, krow0 * ldk + h * HD, 0, S, sKV, AttnSmem<HD>::LDV);
    cp_async_commit();
  }
  for (int rr = 0; rr < TQ / ATT_WARPS; ++rr) {
    const int r = warp * (TQ / ATT_WARPS) + rr, i = i0 + r;
    if (i >= T) {
      if (i < Tm) for (int j = lane; j < ldA; j += 32) dA[(((long long)b * H + h) * Tm + i) * ldA + j] = 0.f;
      continue;
    }
    float* row = sP + (size_t)r * S_pad;
    const long long goff = (((long long)b * H + h) * Tm + i) * ldA;
    float dl = 0.f;
    for (int j = lane; j < ldA; j += 32) {
      const float g = (j < S) ? row[j] : 0.f;
      dA[goff + j] = g * ginv;                // the hooked gradient, unmasked (autograd of bmm(A, v)), in TRUE units
      if (j < S) dl = fmaf(g, A[goff + j], dl);
    }
    dl = warp_sum(dl);
    if (dQ != nullptr) {
      for (int j = lane; j < S; j += 32) row[j] = A[goff + j] * (row[j] - dl);   // dS
    }
    if (lane == 0 && delta) delta[((long long)b * H + h) * Tm + i] = dl;
  }
  if (dQ == nullptr) return;
  __syncthreads();
  float out[HD / 16][4];
  tile_pv<HD, MB>(K, ldk, krow0 * ldk + h * HD, S, sP, S_pad, sKV, out, T - i0);
  store_pv<HD, MB>(dQ, lddq, qrow0, i0, T, h, out, scale);
}

// backward, key side: one CTA = 64 keys of one (b,h) (a whole CLIP ViT-B/32 head); walks the query rows in tiles of 64.
//   dV[j] = sum_i A[i][j] dO[i]      dK[j] = scale * sum_i dS[i][j] Q[i],  dS = A (.) (dA - delta_i)
// Warp w owns the 16-key block w&3 and the d half w>>2 of both products.
constexpr int KV_KEYS = 64, KV_ROWS = 64;
template <int HD>
__global__ void __launch_bounds__(KV_THREADS) attention_bwd_kv_kernel(
    const float* __restrict__ dO, int lddo, const float* __restrict__ Q, int ldq, const float* __restrict__ A,
    const float* __restrict__ dA, int ldA, const float* __restrict__ delta, float* __restrict__ dK, int lddk,
    float* __restrict__ dV, int lddv, int H, int T, int S, float scale, Ragged rg, const float* __restrict__ gscale) {
  constexpr int LDH = HD + 4, LDK = KV_KEYS + 4, NT = HD / 16;   // both are [k][n]-indexed operands: stride = 4 (mod 16)
  extern __shared__ float smem[];
  float* sdO = smem;
  float* sQ = sdO + KV_ROWS * LDH;
  float* sA = sQ + KV_ROWS * LDH;          // [KV_ROWS][LDK]  A[i][j]
  float* sS = sA + KV_ROWS * LDK;          // [KV_ROWS][LDK]  dS[i][j]
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * KV_KEYS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int m0 = (warp & 3) * 16, d0 = (warp >> 2) * (HD / 2);   // this warp's 16 keys and d half
  const int Tm = T;
  long long qrow0 = (long long)b * T, krow0 = (long long)b * S;
  if (rg.lens) { T = S = rg.lens[b]; qrow0 = krow0 = rg.offs[b]; }
  if (j0 >= S) return;
  const float gs = gscale ? gscale[b] : 1.f;           // staged dA is in true units, delta / dO / the outputs are scaled
  float accV[NT][4] = {}, crsV[NT][4] = {}, accK[NT][4] = {}, crsK[NT][4] = {};
  const long long plane = ((long long)b * H + h) * Tm;
  for (int i0 = 0; i0 < T; i0 += KV_ROWS) {
    __syncthreads();
    load_rows_async<HD, KV_ROWS>(dO, lddo, qrow0 * lddo + h * HD, i0, T, sdO, LDH);
    load_rows_async<HD, KV_ROWS>(Q, ldq, qrow0 * ldq + h * HD, i0, T, sQ, LDH);
    cp_async_commit();
    // A and dS tiles: 4 keys per 128-bit load.  Staged rows are ldA (% 4 == 0) wide and zero beyond this sample's
    // keys, so a float4 that starts below ldA is entirely readable and entirely correct.
    for (int e = tid; e < KV_ROWS * (KV_KEYS / 4); e += KV_THREADS) {
      const int r = e / (KV_KEYS / 4), c = (e % (KV_KEYS / 4)) * 4;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), ds = a;
      if (i0 + r < T && j0 + c < ldA) {
        const long long off = (plane + i0 + r) * ldA + j0 + c;
        a = *reinterpret_cast<const float4*>(A + off);
        const float4 ga = *reinterpret_cast<const float4*>(dA + off);
        const float dl = delta[plane + i0 + r];
        ds = make_float4(a.x * (ga.x * gs - dl), a.y * (ga.y * gs - dl), a.z * (ga.z * gs - dl), a.w * (ga.w * gs - dl));
      }
      *reinterpret_cast<float4*>(sA + r * LDK + c) = a;
      *reinterpret_cast<float4*>(sS + r * LDK + c) = ds;
    }
    cp_async_wait_all();
    __syncthreads();
    const int in = min(KV_ROWS, T - i0);                // rows in .. round_up(in, 16) are zero-filled
    if (j0 + m0 < S) {                                  // this warp's 16 keys are live (warp-uniform)
#pragma unroll 2
      for (int kk = 0; kk < in; kk += 16) {
        const FragA aA = frag_a_kmajor(sA + (kk + 2 * t) * LDK + m0 + g, LDK);
        const FragA aS = frag_a_kmajor(sS + (kk + 2 * t) * LDK + m0 + g, LDK);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const FragB bo = frag_b_kmajor(sdO + (kk + 2 * t) * LDH + d0 + nt * 8 + g, LDH);
          mma3(accV[nt], crsV[nt], aA, bo);
          const FragB bq = frag_b_kmajor(sQ + (kk + 2 * t) * LDH + d0 + nt * 8 + g, LDH);
          mma3(accK[nt], crsK[nt], aS, bq);
        }
      }
    }
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int j = j0 + m0 + g + 8 * half;
    if (j >= S) continue;
    float* vrow = dV + (krow0 + j) * lddv + h * HD + d0 + 2 * t;
    float* krow = dK + (krow0 + j) * lddk + h * HD + d0 + 2 * t;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      *reinterpret_cast<float2*>(vrow + nt * 8) = make_float2(fmaf(crsV[nt][2 * half], LO_INV, accV[nt][2 * half]),
                                                              fmaf(crsV[nt][2 * half + 1], LO_INV, accV[nt][2 * half + 1]));
      *reinterpret_cast<float2*>(krow + nt * 8) = make_float2(fmaf(crsK[nt][2 * half], LO_INV, accK[nt][2 * half]) * scale,
                                                              fmaf(crsK[nt][2 * half + 1], LO_INV, accK[nt][2 * half + 1]) * scale);
    }
  }
}

constexpr size_t ATT_SMEM_MAX = 227 * 1024;

template <int HD, int TQ>
static int launch_fwd_tq(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const float* key_bias,
                         float* A, int ldA, float* O, int ldo, int B, int H, int T, int S, float scale, int flags,
                         Ragged rg, cudaStream_t st) {
  const size_t smem = AttnSmem<HD>::bytes(S, TQ);
  MMX_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<HD, TQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(T, TQ), H, B);
  attention_fwd_kernel<HD, TQ><<<grid, TQ * 4, smem, st>>>(Q, ldq, K, ldk, V, ldv, key_bias, A, ldA, O, ldo, H, T, S, scale,
                                                          flags, rg);
  MMX_LAUNCH_CHECK();
  return 0;
}

template <int HD>
static int launch_fwd(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const float* key_bias,
                      float* A, int ldA, float* O, int ldo, int B, int H, int T, int S, float scale, int flags,
                      Ragged rg, cudaStream_t st) {
  if (AttnSmem<HD>::bytes(S, 64) <= ATT_SMEM_MAX)
    return launch_fwd_tq<HD, 64>(Q, ldq, K, ldk, V, ldv, key_bias, A, ldA, O, ldo, B, H, T, S, scale, flags, rg, st);
  MMX_REQUIRE(AttnSmem<HD>::bytes(S, 32) <= ATT_SMEM_MAX, "sequence too long for the single-pass attention kernel");
  return launch_fwd_tq<HD, 32>(Q, ldq, K, ldk, V, ldv, key_bias, A, ldA, O, ldo, B, H, T, S, scale, flags, rg, st);
}

template <int HD, int TQ>
static int launch_bwd_q_tq(const float* dO, int lddo, const float* K, int ldk, const float* V, int ldv, const float* A,
                           float* dA, int ldA, float* delta, float* dQ, int lddq, int B, int H, int T, int S, float scale,
                           Ragged rg, const float* gscale, cudaStream_t st) {
  const size_t smem = AttnSmem<HD>::bytes(S, TQ);
  MMX_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_q_kernel<HD, TQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(T, TQ), H, B);
  attention_bwd_q_kernel<HD, TQ><<<grid, TQ * 4, smem, st>>>(dO, lddo, K, ldk, V, ldv, A, dA, ldA, delta, dQ, lddq, H, T, S,
                                                            scale, rg, gscale);
  MMX_LAUNCH_CHECK();
  return 0;
}

template <int HD>
static int launch_bwd(const float* dO, int lddo, const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv,
                      const float* A, float* dA, int ldA, float* delta, float* dQ, int lddq, float* dK, int lddk, float* dV,
                      int lddv, int B, int H, int T, int S, float scale, Ragged rg, const float* gscale, cudaStream_t st) {
  if (AttnSmem<HD>::bytes(S, 64) <= ATT_SMEM_MAX) {
    MMX_TRY((launch_bwd_q_tq<HD, 64>(dO, lddo, K, ldk, V, ldv, A, dA, ldA, delta, dQ, lddq, B, H, T, S, scale, rg, gscale, st)));
  } else {
    MMX_REQUIRE(AttnSmem<HD>::bytes(S, 32) <= ATT_SMEM_MAX, "sequence too long for the single-pass attention kernel");
    MMX_TRY((launch_bwd_q_tq<HD, 32>(dO, lddo, K, ldk, V, ldv, A, dA, ldA, delta, dQ, lddq, B, H, T, S, scale, rg, gscale, st)));
  }
  if (dQ == nullptr) return 0;
  const size_t smem2 = sizeof(float) * (2 * KV_ROWS * (HD + 4) + 2 * KV_ROWS * (KV_KEYS + 4));
  MMX_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  dim3 grid2(cdiv(S, KV_KEYS), H, B);
  attention_bwd_kv_kernel<HD><<<grid2, KV_THREADS, smem2, st>>>(dO, lddo, Q, ldq, A, dA, ldA, delta, dK, lddk, dV, lddv, H,
                                                                T, S, scale, rg, gscale);
  MMX_LAUNCH_CHECK();
  return 0;
}

int attention_fwd(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const float* key_bias, float* A,
                  int ldA, float* O, int ldo, int B, int H, int T, int S, int hd, float scale, int flags, cudaStream_t st,
                  const int* offs, const int* lens) {
  Ragged rg;
  rg.offs = offs; rg.lens = lens;
  MMX_REQUIRE(lens == nullptr || (T == S && key_bias == nullptr && offs != nullptr), "ragged attention: self-attention without key bias only");
  MMX_REQUIRE(ldA >= S && ldA % 4 == 0, "ldA must be >= S and a multiple of 4");
  MMX_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0 && aligned16(Q) && aligned16(K) && aligned16(V) &&
                  aligned16(O) && aligned16(A), "attention operands must be 16-byte aligned with leading dims % 4 == 0");
  if (B == 0 || T == 0) return 0;
  switch (hd) {
    case 16: return launch_fwd<16>(Q, ldq, K, ldk, V, ldv, key_bias, A, ldA, O, ldo, B, H, T, S, scale, flags, rg, st);
    case 32: return launch_fwd<32>(Q, ldq, K, ldk, V, ldv, key_bias, A, ldA, O, ldo, B, H, T, S, scale, flags, rg, st);
    case 64: return launch_fwd<64>(Q, ldq, K, ldk, V, ldv, key_bias, A, ldA, O, ldo, B, H, T, S, scale, flags, rg, st);
    default: MMX_REQUIRE(false, "head_dim must be 16, 32 or 64");
  }
  return 0;
}

int attention_bwd(const float* dO, int lddo, const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv,
                  const float* A, float* dA, int ldA, float* delta, float* dQ, int lddq, float* dK, int lddk, float* dV,
                  int lddv, int B, int H, int T, int S, int hd, float scale, int flags, cudaStream_t st, const int* offs,
                  const int* lens, const float* gscale) {
  Ragged rg;
  rg.offs = offs; rg.lens = lens;
  MMX_REQUIRE(lens == nullptr || (T == S && offs != nullptr), "ragged attention: self-attention only");
  MMX_REQUIRE(ldA >= S && ldA % 4 == 0, "ldA must be >= S and a multiple of 4");
  MMX_REQUIRE(lddo % 4 == 0 && ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && aligned16(dO) && aligned16(Q) && aligned16(K) &&
                  aligned16(V) && aligned16(A) && aligned16(dA) && (dQ == nullptr || (lddq % 4 == 0 && lddk % 4 == 0 &&
                  lddv % 4 == 0 && aligned16(dQ) && aligned16(dK) && aligned16(dV))),
              "attention operands must be 16-byte aligned with leading dims % 4 == 0");
  MMX_REQUIRE((dQ == nullptr) == (dK == nullptr) && (dQ == nullptr) == (dV == nullptr), "dQ/dK/dV: all or none");
  MMX_REQUIRE(dQ == nullptr || delta != nullptr, "delta scratch required");
  (void)flags;  // the mask is already folded into A (masked entries are exactly 0, so dS vanishes there)
  if (B == 0 || T == 0) return 0;
  switch (hd) {
    case 16: return launch_bwd<16>(dO, lddo, Q, ldq, K, ldk, V, ldv, A, dA, ldA, delta, dQ, lddq, dK, lddk, dV, lddv, B, H, T, S, scale, rg, gscale, st);
    case 32: return launch_bwd<32>(dO, lddo, Q, ldq, K, ldk, V, ldv, A, dA, ldA, delta, dQ, lddq, dK, lddk, dV, lddv, B, H, T, S, scale, rg, gscale, st);
    case 64: return launch_bwd<64>(dO, lddo, Q, ldq, K, ldk, V, ldv, A, dA, ldA, delta, dQ, lddq, dK, lddk, dV, lddv, B, H, T, S, scale, rg, gscale, st);
    default: MMX_REQUIRE(false, "head_dim must be 16, 32 or 64");
  }
  return 0;
}

}  // namespace mmx

using namespace mmx;
extern "C" {
int mmx_attention_fwd(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const float* key_bias,
                      float* A, int ldA, float* O, int ldo, int B, int H, int T, int S, int hd, float scale, int flags,
                      void* stream) {
  return attention_fwd(Q, ldq, K, ldk, V, ldv, key_bias, A, ldA, O, ldo, B, H, T, S, hd, scale, flags, (cudaStream_t)stream,
                       nullptr, nullptr);
}
int mmx_attention_bwd(const float* dO, int lddo, const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv,
                      const float* A, float* dA, int ldA, float* delta, float* dQ, int lddq, float* dK, int lddk, float* dV,
                      int lddv, int B, int H, int T, int S, int hd, float scale, int flags, void* stream) {
  return attention_bwd(dO, lddo, Q, ldq, K, ldk, V, ldv, A, dA, ldA, delta, dQ, lddq, dK, lddk, dV, lddv, B, H, T, S, hd, scale,
                       flags, (cudaStream_t)stream, nullptr, nullptr, nullptr);
}
int mmx_attention_bwd_scaled(const float* dO, int lddo, const float* Q, int ldq, const float* K, int ldk, const float* V,
                             int ldv, const float* A, float* dA, int ldA, float* delta, float* dQ, int lddq, float* dK,
                             int lddk, float* dV, int lddv, int B, int H, int T, int S, int hd, float scale, int flags,
                             const float* gscale, void* stream) {
  return attention_bwd(dO, lddo, Q, ldq, K, ldk, V, ldv, A, dA, ldA, delta, dQ, lddq, dK, lddk, dV, lddv, B, H, T, S, hd, scale,
                       flags, (cudaStream_t)stream, nullptr, nullptr, gscale);
}
}
