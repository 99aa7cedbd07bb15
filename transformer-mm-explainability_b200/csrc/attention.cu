// Attention forward that stages A = softmax(QK^T) to HBM, and the attention backward that stages dA = dO V^T
// (the two tensors the reference captures with forward / backward hooks) and continues to dQ, dK, dV.
// fp32 FFMA; one CTA = one (batch, head, 32-query tile); score rows live in shared memory so any S <= ~1500
// (DETR 850, ViT-L/14@336 577) is handled without a second pass.  Deterministic (no atomics): dK/dV come from a
// second kernel that walks the query tiles for one key tile.
#include "mmx_common.cuh"
#include <math_constants.h>

namespace mmx {

// Ragged self-attention (packed rows): sample b owns rows offs[b] .. offs[b]+lens[b]-1 of Q/K/V/O and attends within
// them; the staged A / dA planes stay dense [B,H,T,ld] with rows and columns >= lens[b] zero-filled.
struct Ragged {
  const int* offs = nullptr;
  const int* lens = nullptr;
};

constexpr int TQ = 32;    // query rows per CTA
constexpr int TKEY = 64;  // keys per shared-memory tile
constexpr int ATT_THREADS = 128;

template <int HD>
struct AttnSmem {
  static constexpr int LDH = HD + 4;
  static size_t bytes(int S_pad) { return sizeof(float) * ((size_t)TQ * S_pad + (size_t)TQ * LDH + (size_t)TKEY * LDH); }
};

// Shared-memory operand rows are HD + 4 floats: 16-byte aligned for 128-bit reads, and 8 consecutive rows start in 8
// different 4-bank groups, so a quarter-warp's LDS.128 is conflict-free.
//
// scores[i][j] (i in tile, j in [0,S)) = sum_d X[i][d] * Y[j][d];  X rows already in sX, Y streamed through sY.
template <int HD>
__device__ __forceinline__ void tile_scores(const float* __restrict__ Y, int ldy, long long ybase, int S,
                                            const float* sX, float* sY, float* sP, int S_pad, float post_scale) {
  constexpr int LDH = HD + 4;
  const int tid = threadIdx.x, ri = tid >> 4, cj = tid & 15;
  for (int j0 = 0; j0 < S; j0 += TKEY) {
    __syncthreads();
    for (int e = tid; e < TKEY * (HD / 4); e += ATT_THREADS) {
      const int r = e / (HD / 4), d = (e % (HD / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + r < S) v = *reinterpret_cast<const float4*>(Y + ybase + (long long)(j0 + r) * ldy + d);
      *reinterpret_cast<float4*>(sY + r * LDH + d) = v;
    }
    __syncthreads();
    float acc[4][4] = {};
#pragma unroll 4
    for (int d = 0; d < HD; d += 4) {
      float4 x[4], y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        x[u] = *reinterpret_cast<const float4*>(sX + (ri * 4 + u) * LDH + d);
        y[u] = *reinterpret_cast<const float4*>(sY + (cj + 16 * u) * LDH + d);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc[a][c] = fmaf(x[a].x, y[c].x, acc[a][c]);
          acc[a][c] = fmaf(x[a].y, y[c].y, acc[a][c]);
          acc[a][c] = fmaf(x[a].z, y[c].z, acc[a][c]);
          acc[a][c] = fmaf(x[a].w, y[c].w, acc[a][c]);
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + cj + 16 * c;
        if (j < S_pad) sP[(ri * 4 + a) * S_pad + j] = acc[a][c] * post_scale;
      }
  }
  __syncthreads();
}

// out[i][d] = sum_j sP[i][j] * Y[j][d];  thread (ri, cj) owns rows ri*4..+3 and the 4 contiguous columns d = 4*cj..
// (threads with 4*cj >= HD idle for small head dims)
template <int HD>
__device__ __forceinline__ void tile_pv(const float* __restrict__ Y, int ldy, long long ybase, int S, const float* sP,
                                        int S_pad, float* sY, float4 (&out)[4]) {
  constexpr int LDH = HD + 4;
  const int tid = threadIdx.x, ri = tid >> 4, cj = tid & 15;
  const bool active = cj * 4 < HD;
#pragma unroll
  for (int a = 0; a < 4; ++a) out[a] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j0 = 0; j0 < S; j0 += TKEY) {
    __syncthreads();
    for (int e = tid; e < TKEY * (HD / 4); e += ATT_THREADS) {
      const int r = e / (HD / 4), d = (e % (HD / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + r < S) v = *reinterpret_cast<const float4*>(Y + ybase + (long long)(j0 + r) * ldy + d);
      *reinterpret_cast<float4*>(sY + r * LDH + d) = v;
    }
    __syncthreads();
    if (!active) continue;
    const int jn = min(TKEY, S - j0);
    int j = 0;
    for (; j + 4 <= jn; j += 4) {                       // S_pad % 4 == 0 and j0 % 4 == 0: aligned float4 reads of P
      float4 p[4], y[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) p[a] = *reinterpret_cast<const float4*>(sP + (ri * 4 + a) * S_pad + j0 + j);
#pragma unroll
      for (int u = 0; u < 4; ++u) y[u] = *reinterpret_cast<const float4*>(sY + (j + u) * LDH + cj * 4);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float pa[4] = {p[a].x, p[a].y, p[a].z, p[a].w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          out[a].x = fmaf(pa[u], y[u].x, out[a].x);
          out[a].y = fmaf(pa[u], y[u].y, out[a].y);
          out[a].z = fmaf(pa[u], y[u].z, out[a].z);
          out[a].w = fmaf(pa[u], y[u].w, out[a].w);
        }
      }
    }
    for (; j < jn; ++j) {
      const float4 y = *reinterpret_cast<const float4*>(sY + j * LDH + cj * 4);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float pa = sP[(ri * 4 + a) * S_pad + j0 + j];
        out[a].x = fmaf(pa, y.x, out[a].x); out[a].y = fmaf(pa, y.y, out[a].y);
        out[a].z = fmaf(pa, y.z, out[a].z); out[a].w = fmaf(pa, y.w, out[a].w);
      }
    }
  }
}

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attention_fwd_kernel(
    const float* __restrict__ Q, int ldq, const float* __restrict__ K, int ldk, const float* __restrict__ V, int ldv,
    const float* __restrict__ key_bias, float* __restrict__ A, int ldA, float* __restrict__ O, int ldo, int H, int T, int S,
    float scale, int flags, Ragged rg) {
  constexpr int LDH = HD + 4;
  extern __shared__ float smem[];
  const int S_pad = ldA;  // score rows use the same padded width as the staged A rows
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
  const float qs = scale_scores ? 1.f : scale;
  for (int e = tid; e < TQ * (HD / 4); e += ATT_THREADS) {
    const int r = e / (HD / 4), d = (e % (HD / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i0 + r < T) v = *reinterpret_cast<const float4*>(Q + (qrow0 + i0 + r) * ldq + h * HD + d);
    v.x *= qs; v.y *= qs; v.z *= qs; v.w *= qs;
    *reinterpret_cast<float4*>(sQ + r * LDH + d) = v;
  }
  tile_scores<HD>(K, ldk, krow0 * ldk + h * HD, S, sQ, sKV, sP, S_pad, scale_scores ? scale : 1.f);
  // softmax per row (warp w owns rows w*8 .. w*8+7), stage A
  for (int rr = 0; rr < TQ / 4; ++rr) {
    const int r = warp * (TQ / 4) + rr, i = i0 + r;
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
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) { const float e = expf(row[j] - mx); row[j] = e; sum += e; }
    sum = warp_sum(sum);
    float* arow = A + (((long long)b * H + h) * Tm + i) * ldA;
    for (int j = lane; j < ldA; j += 32) {
      float p = 0.f;
      if (j < S) { p = row[j] / sum; row[j] = p; }
      arow[j] = p;
    }
  }
  __syncthreads();
  float4 out[4];
  tile_pv<HD>(V, ldv, krow0 * ldv + h * HD, S, sP, S_pad, sKV, out);
  const int ri = tid >> 4, cj = tid & 15;
  if (cj * 4 < HD) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = i0 + ri * 4 + a;
      if (i < T) *reinterpret_cast<float4*>(O + (qrow0 + i) * ldo + h * HD + cj * 4) = out[a];
    }
  }
}

// backward, query side: dA (staged), delta, dQ
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attention_bwd_q_kernel(
    const float* __restrict__ dO, int lddo, const float* __restrict__ K, int ldk, const float* __restrict__ V, int ldv,
    const float* __restrict__ A, float* __restrict__ dA, int ldA, float* __restrict__ delta, float* __restrict__ dQ,
    int lddq, int H, int T, int S, float scale, Ragged rg) {
  constexpr int LDH = HD + 4;
  extern __shared__ float smem[];
  const int S_pad = ldA;
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
  for (int e = tid; e < TQ * (HD / 4); e += ATT_THREADS) {
    const int r = e / (HD / 4), d = (e % (HD / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i0 + r < T) v = *reinterpret_cast<const float4*>(dO + (qrow0 + i0 + r) * lddo + h * HD + d);
    *reinterpret_cast<float4*>(sX + r * LDH + d) = v;
  }
  tile_scores<HD>(V, ldv, krow0 * ldv + h * HD, S, sX, sKV, sP, S_pad, 1.f);
  for (int rr = 0; rr < TQ / 4; ++rr) {
    const int r = warp * (TQ / 4) + rr, i = i0 + r;
    if (i >= T) {
      if (i < Tm) for (int j = lane; j < ldA; j += 32) dA[(((long long)b * H + h) * Tm + i) * ldA + j] = 0.f;
      continue;
    }
    float* row = sP + (size_t)r * S_pad;
    const long long goff = (((long long)b * H + h) * Tm + i) * ldA;
    float dl = 0.f;
    for (int j = lane; j < ldA; j += 32) {
      const float g = (j < S) ? row[j] : 0.f;
      dA[goff + j] = g;                       // the hooked gradient, unmasked (autograd of bmm(A, v))
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
  float4 out[4];
  tile_pv<HD>(K, ldk, krow0 * ldk + h * HD, S, sP, S_pad, sKV, out);
  const int ri = tid >> 4, cj = tid & 15;
  if (cj * 4 < HD) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = i0 + ri * 4 + a;
      if (i < T) {
        const float4 v = make_float4(out[a].x * scale, out[a].y * scale, out[a].z * scale, out[a].w * scale);
        *reinterpret_cast<float4*>(dQ + (qrow0 + i) * lddq + h * HD + cj * 4) = v;
      }
    }
  }
}

// backward, key side: one CTA = 32 keys of one (b,h); walks all query rows in tiles of 64.
//   dV[j] = sum_i A[i][j] dO[i]      dK[j] = scale * sum_i dS[i][j] Q[i],  dS = A (.) (dA - delta_i)
constexpr int KV_KEYS = 32, KV_ROWS = 64;
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attention_bwd_kv_kernel(
    const float* __restrict__ dO, int lddo, const float* __restrict__ Q, int ldq, const float* __restrict__ A,
    const float* __restrict__ dA, int ldA, const float* __restrict__ delta, float* __restrict__ dK, int lddk,
    float* __restrict__ dV, int lddv, int H, int T, int S, float scale, Ragged rg) {
  constexpr int LDH = HD + 4, LDK = KV_KEYS + 4;
  extern __shared__ float smem[];
  float* sdO = smem;
  float* sQ = sdO + KV_ROWS * LDH;
  float* sA = sQ + KV_ROWS * LDH;          // [KV_ROWS][LDK]  A[i][j]
  float* sS = sA + KV_ROWS * LDK;          // [KV_ROWS][LDK]  dS[i][j]
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * KV_KEYS;
  const int tid = threadIdx.x, rj = tid >> 4, cj = tid & 15;
  const int Tm = T;
  long long qrow0 = (long long)b * T, krow0 = (long long)b * S;
  if (rg.lens) { T = S = rg.lens[b]; qrow0 = krow0 = rg.offs[b]; }
  if (j0 >= S) return;
  const bool active = cj * 4 < HD;
  float4 accV[4], accK[4];
#pragma unroll
  for (int x = 0; x < 4; ++x) accV[x] = accK[x] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long plane = ((long long)b * H + h) * Tm;
  for (int i0 = 0; i0 < T; i0 += KV_ROWS) {
    __syncthreads();
    for (int e = tid; e < KV_ROWS * KV_KEYS; e += ATT_THREADS) {
      const int r = e / KV_KEYS, c = e % KV_KEYS;
      float a = 0.f, ds = 0.f;
      if (i0 + r < T && j0 + c < S) {
        const long long off = (plane + i0 + r) * ldA + j0 + c;
        a = A[off];
        ds = a * (dA[off] - delta[plane + i0 + r]);
      }
      sA[r * LDK + c] = a;
      sS[r * LDK + c] = ds;
    }
    for (int e = tid; e < KV_ROWS * (HD / 4); e += ATT_THREADS) {
      const int r = e / (HD / 4), d = (e % (HD / 4)) * 4;
      float4 vo = make_float4(0.f, 0.f, 0.f, 0.f), vq = vo;
      if (i0 + r < T) {
        vo = *reinterpret_cast<const float4*>(dO + (qrow0 + i0 + r) * lddo + h * HD + d);
        vq = *reinterpret_cast<const float4*>(Q + (qrow0 + i0 + r) * ldq + h * HD + d);
      }
      *reinterpret_cast<float4*>(sdO + r * LDH + d) = vo;
      *reinterpret_cast<float4*>(sQ + r * LDH + d) = vq;
    }
    __syncthreads();
    if (!active) continue;
    const int in = min(KV_ROWS, T - i0);
    for (int i = 0; i < in; ++i) {
      const float4 a = *reinterpret_cast<const float4*>(sA + i * LDK + rj * 4);
      const float4 sv = *reinterpret_cast<const float4*>(sS + i * LDK + rj * 4);
      const float4 o = *reinterpret_cast<const float4*>(sdO + i * LDH + cj * 4);
      const float4 q = *reinterpret_cast<const float4*>(sQ + i * LDH + cj * 4);
      const float av[4] = {a.x, a.y, a.z, a.w}, ss[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        accV[x].x = fmaf(av[x], o.x, accV[x].x); accV[x].y = fmaf(av[x], o.y, accV[x].y);
        accV[x].z = fmaf(av[x], o.z, accV[x].z); accV[x].w = fmaf(av[x], o.w, accV[x].w);
        accK[x].x = fmaf(ss[x], q.x, accK[x].x); accK[x].y = fmaf(ss[x], q.y, accK[x].y);
        accK[x].z = fmaf(ss[x], q.z, accK[x].z); accK[x].w = fmaf(ss[x], q.w, accK[x].w);
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int j = j0 + rj * 4 + x;
    if (j >= S) continue;
    *reinterpret_cast<float4*>(dV + (krow0 + j) * lddv + h * HD + cj * 4) = accV[x];
    const float4 kk = make_float4(accK[x].x * scale, accK[x].y * scale, accK[x].z * scale, accK[x].w * scale);
    *reinterpret_cast<float4*>(dK + (krow0 + j) * lddk + h * HD + cj * 4) = kk;
  }
}

template <int HD>
static int launch_fwd(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const float* key_bias,
                      float* A, int ldA, float* O, int ldo, int B, int H, int T, int S, float scale, int flags,
                      Ragged rg, cudaStream_t st) {
  const size_t smem = AttnSmem<HD>::bytes(ldA);
  MMX_REQUIRE(smem <= 227 * 1024, "sequence too long for the single-pass attention kernel");
  MMX_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(T, TQ), H, B);
  attention_fwd_kernel<HD><<<grid, ATT_THREADS, smem, st>>>(Q, ldq, K, ldk, V, ldv, key_bias, A, ldA, O, ldo, H, T, S, scale,
                                                            flags, rg);
  MMX_LAUNCH_CHECK();
  return 0;
}

template <int HD>
static int launch_bwd(const float* dO, int lddo, const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv,
                      const float* A, float* dA, int ldA, float* delta, float* dQ, int lddq, float* dK, int lddk, float* dV,
                      int lddv, int B, int H, int T, int S, float scale, Ragged rg, cudaStream_t st) {
  const size_t smem = AttnSmem<HD>::bytes(ldA);
  MMX_REQUIRE(smem <= 227 * 1024, "sequence too long for the single-pass attention kernel");
  MMX_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_q_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(T, TQ), H, B);
  attention_bwd_q_kernel<HD><<<grid, ATT_THREADS, smem, st>>>(dO, lddo, K, ldk, V, ldv, A, dA, ldA, delta, dQ, lddq, H, T, S,
                                                              scale, rg);
  MMX_LAUNCH_CHECK();
  if (dQ == nullptr) return 0;
  const size_t smem2 = sizeof(float) * (2 * KV_ROWS * (HD + 4) + 2 * KV_ROWS * (KV_KEYS + 4));
  MMX_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  dim3 grid2(cdiv(S, KV_KEYS), H, B);
  attention_bwd_kv_kernel<HD><<<grid2, ATT_THREADS, smem2, st>>>(dO, lddo, Q, ldq, A, dA, ldA, delta, dK, lddk, dV, lddv, H,
                                                                 T, S, scale, rg);
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
                  const int* lens) {
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
    case 16: return launch_bwd<16>(dO, lddo, Q, ldq, K, ldk, V, ldv, A, dA, ldA, delta, dQ, lddq, dK, lddk, dV, lddv, B, H, T, S, scale, rg, st);
    case 32: return launch_bwd<32>(dO, lddo, Q, ldq, K, ldk, V, ldv, A, dA, ldA, delta, dQ, lddq, dK, lddk, dV, lddv, B, H, T, S, scale, rg, st);
    case 64: return launch_bwd<64>(dO, lddo, Q, ldq, K, ldk, V, ldv, A, dA, ldA, delta, dQ, lddq, dK, lddk, dV, lddv, B, H, T, S, scale, rg, st);
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
                       flags, (cudaStream_t)stream, nullptr, nullptr);
}
}
