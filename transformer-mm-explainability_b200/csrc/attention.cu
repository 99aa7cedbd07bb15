// Attention forward that stages A = softmax(QK^T) to HBM, and the attention backward that stages dA = dO V^T
// (the two tensors the reference captures with forward / backward hooks) and continues to dQ, dK, dV.
// fp32 FFMA; one CTA = one (batch, head, 32-query tile); score rows live in shared memory so any S <= ~1500
// (DETR 850, ViT-L/14@336 577) is handled without a second pass.  Deterministic (no atomics): dK/dV come from a
// second kernel that walks the query tiles for one key tile.
#include "mmx_common.cuh"
#include <math_constants.h>

namespace mmx {

constexpr int TQ = 32;    // query rows per CTA
constexpr int TKEY = 64;  // keys per shared-memory tile
constexpr int ATT_THREADS = 128;

template <int HD>
struct AttnSmem {
  static constexpr int LDH = HD + 1;
  static size_t bytes(int S_pad) { return sizeof(float) * ((size_t)TQ * S_pad + (size_t)TQ * LDH + (size_t)TKEY * LDH); }
};

// scores[i][j] (i in tile, j in [0,S)) = sum_d X[i][d] * Y[j][d];  X rows already in sX, Y streamed through sY.
template <int HD>
__device__ __forceinline__ void tile_scores(const float* __restrict__ Y, int ldy, long long ybase, int S,
                                            const float* sX, float* sY, float* sP, int S_pad, float post_scale) {
  constexpr int LDH = HD + 1;
  const int tid = threadIdx.x, ri = tid >> 4, cj = tid & 15;
  for (int j0 = 0; j0 < S; j0 += TKEY) {
    __syncthreads();
    for (int e = tid; e < TKEY * HD; e += ATT_THREADS) {
      const int r = e / HD, d = e % HD;
      sY[r * LDH + d] = (j0 + r < S) ? Y[ybase + (long long)(j0 + r) * ldy + d] : 0.f;
    }
    __syncthreads();
    float acc[4][4] = {};
#pragma unroll 8
    for (int d = 0; d < HD; ++d) {
      float x[4], y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { x[u] = sX[(ri * 4 + u) * LDH + d]; y[u] = sY[(cj + 16 * u) * LDH + d]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(x[a], y[c], acc[a][c]);
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

// out[i][d] = sum_j sP[i][j] * Y[j][d]
template <int HD>
__device__ __forceinline__ void tile_pv(const float* __restrict__ Y, int ldy, long long ybase, int S, const float* sP,
                                        int S_pad, float* sY, float (&out)[4][HD / 16]) {
  constexpr int LDH = HD + 1, NU = HD / 16;
  const int tid = threadIdx.x, ri = tid >> 4, cj = tid & 15;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int u = 0; u < NU; ++u) out[a][u] = 0.f;
  for (int j0 = 0; j0 < S; j0 += TKEY) {
    __syncthreads();
    for (int e = tid; e < TKEY * HD; e += ATT_THREADS) {
      const int r = e / HD, d = e % HD;
      sY[r * LDH + d] = (j0 + r < S) ? Y[ybase + (long long)(j0 + r) * ldy + d] : 0.f;
    }
    __syncthreads();
    const int jn = min(TKEY, S - j0);
    for (int j = 0; j < jn; ++j) {
      float p[4], y[NU];
#pragma unroll
      for (int a = 0; a < 4; ++a) p[a] = sP[(ri * 4 + a) * S_pad + j0 + j];
#pragma unroll
      for (int u = 0; u < NU; ++u) y[u] = sY[j * LDH + cj + 16 * u];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int u = 0; u < NU; ++u) out[a][u] = fmaf(p[a], y[u], out[a][u]);
    }
  }
}

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attention_fwd_kernel(
    const float* __restrict__ Q, int ldq, const float* __restrict__ K, int ldk, const float* __restrict__ V, int ldv,
    const float* __restrict__ key_bias, float* __restrict__ A, int ldA, float* __restrict__ O, int ldo, int H, int T, int S,
    float scale, int flags) {
  constexpr int LDH = HD + 1, NU = HD / 16;
  extern __shared__ float smem[];
  const int S_pad = ldA;  // score rows use the same padded width as the staged A rows
  float* sP = smem;
  float* sQ = sP + (size_t)TQ * S_pad;
  float* sKV = sQ + TQ * LDH;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * TQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool scale_scores = flags & MMX_ATTN_SCALE_SCORES;
  const float qs = scale_scores ? 1.f : scale;
  for (int e = tid; e < TQ * HD; e += ATT_THREADS) {
    const int r = e / HD, d = e % HD;
    sQ[r * LDH + d] = (i0 + r < T) ? Q[((long long)b * T + i0 + r) * ldq + h * HD + d] * qs : 0.f;
  }
  tile_scores<HD>(K, ldk, (long long)b * S * ldk + h * HD, S, sQ, sKV, sP, S_pad, scale_scores ? scale : 1.f);
  // softmax per row (warp w owns rows w*8 .. w*8+7), stage A
  for (int rr = 0; rr < TQ / 4; ++rr) {
    const int r = warp * (TQ / 4) + rr, i = i0 + r;
    if (i >= T) continue;
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
    float* arow = A + (((long long)b * H + h) * T + i) * ldA;
    for (int j = lane; j < ldA; j += 32) {
      float p = 0.f;
      if (j < S) { p = row[j] / sum; row[j] = p; }
      arow[j] = p;
    }
  }
  __syncthreads();
  float out[4][NU];
  tile_pv<HD>(V, ldv, (long long)b * S * ldv + h * HD, S, sP, S_pad, sKV, out);
  const int ri = tid >> 4, cj = tid & 15;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ri * 4 + a;
    if (i >= T) continue;
#pragma unroll
    for (int u = 0; u < NU; ++u) O[((long long)b * T + i) * ldo + h * HD + cj + 16 * u] = out[a][u];
  }
}

// backward, query side: dA (staged), delta, dQ
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attention_bwd_q_kernel(
    const float* __restrict__ dO, int lddo, const float* __restrict__ K, int ldk, const float* __restrict__ V, int ldv,
    const float* __restrict__ A, float* __restrict__ dA, int ldA, float* __restrict__ delta, float* __restrict__ dQ,
    int lddq, int H, int T, int S, float scale) {
  constexpr int LDH = HD + 1, NU = HD / 16;
  extern __shared__ float smem[];
  const int S_pad = ldA;
  float* sP = smem;
  float* sX = sP + (size_t)TQ * S_pad;
  float* sKV = sX + TQ * LDH;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * TQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < TQ * HD; e += ATT_THREADS) {
    const int r = e / HD, d = e % HD;
    sX[r * LDH + d] = (i0 + r < T) ? dO[((long long)b * T + i0 + r) * lddo + h * HD + d] : 0.f;
  }
  tile_scores<HD>(V, ldv, (long long)b * S * ldv + h * HD, S, sX, sKV, sP, S_pad, 1.f);
  for (int rr = 0; rr < TQ / 4; ++rr) {
    const int r = warp * (TQ / 4) + rr, i = i0 + r;
    if (i >= T) continue;
    float* row = sP + (size_t)r * S_pad;
    const long long goff = (((long long)b * H + h) * T + i) * ldA;
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
    if (lane == 0 && delta) delta[((long long)b * H + h) * T + i] = dl;
  }
  if (dQ == nullptr) return;
  __syncthreads();
  float out[4][NU];
  tile_pv<HD>(K, ldk, (long long)b * S * ldk + h * HD, S, sP, S_pad, sKV, out);
  const int ri = tid >> 4, cj = tid & 15;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ri * 4 + a;
    if (i >= T) continue;
#pragma unroll
    for (int u = 0; u < NU; ++u) dQ[((long long)b * T + i) * lddq + h * HD + cj + 16 * u] = out[a][u] * scale;
  }
}

// backward, key side: one CTA = 32 keys of one (b,h); walks all query rows in tiles of 64.
//   dV[j] = sum_i A[i][j] dO[i]      dK[j] = scale * sum_i dS[i][j] Q[i],  dS = A (.) (dA - delta_i)
constexpr int KV_KEYS = 32, KV_ROWS = 64;
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attention_bwd_kv_kernel(
    const float* __restrict__ dO, int lddo, const float* __restrict__ Q, int ldq, const float* __restrict__ A,
    const float* __restrict__ dA, int ldA, const float* __restrict__ delta, float* __restrict__ dK, int lddk,
    float* __restrict__ dV, int lddv, int H, int T, int S, float scale) {
  constexpr int LDH = HD + 1, NU = HD / 16;
  extern __shared__ float smem[];
  float* sdO = smem;
  float* sQ = sdO + KV_ROWS * LDH;
  float (*sA)[KV_KEYS + 1] = reinterpret_cast<float (*)[KV_KEYS + 1]>(sQ + KV_ROWS * LDH);
  float (*sS)[KV_KEYS + 1] = sA + KV_ROWS;
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * KV_KEYS;
  const int tid = threadIdx.x, rj = tid >> 4, cj = tid & 15;
  float accV[4][NU] = {}, accK[4][NU] = {};
  const long long plane = ((long long)b * H + h) * T;
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
      sA[r][c] = a;
      sS[r][c] = ds;
    }
    for (int e = tid; e < KV_ROWS * HD; e += ATT_THREADS) {
      const int r = e / HD, d = e % HD;
      const bool ok = i0 + r < T;
      sdO[r * LDH + d] = ok ? dO[((long long)b * T + i0 + r) * lddo + h * HD + d] : 0.f;
      sQ[r * LDH + d] = ok ? Q[((long long)b * T + i0 + r) * ldq + h * HD + d] : 0.f;
    }
    __syncthreads();
    const int in = min(KV_ROWS, T - i0);
    for (int i = 0; i < in; ++i) {
      float a[4], s[4], o[NU], q[NU];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = sA[i][rj * 4 + u]; s[u] = sS[i][rj * 4 + u]; }
#pragma unroll
      for (int u = 0; u < NU; ++u) { o[u] = sdO[i * LDH + cj + 16 * u]; q[u] = sQ[i * LDH + cj + 16 * u]; }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int u = 0; u < NU; ++u) {
          accV[x][u] = fmaf(a[x], o[u], accV[x][u]);
          accK[x][u] = fmaf(s[x], q[u], accK[x][u]);
        }
    }
  }
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int j = j0 + rj * 4 + x;
    if (j >= S) continue;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      dV[((long long)b * S + j) * lddv + h * HD + cj + 16 * u] = accV[x][u];
      dK[((long long)b * S + j) * lddk + h * HD + cj + 16 * u] = accK[x][u] * scale;
    }
  }
}

template <int HD>
static int launch_fwd(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const float* key_bias,
                      float* A, int ldA, float* O, int ldo, int B, int H, int T, int S, float scale, int flags,
                      cudaStream_t st) {
  const size_t smem = AttnSmem<HD>::bytes(ldA);
  MMX_REQUIRE(smem <= 227 * 1024, "sequence too long for the single-pass attention kernel");
  MMX_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(T, TQ), H, B);
  attention_fwd_kernel<HD><<<grid, ATT_THREADS, smem, st>>>(Q, ldq, K, ldk, V, ldv, key_bias, A, ldA, O, ldo, H, T, S, scale,
                                                            flags);
  MMX_LAUNCH_CHECK();
  return 0;
}

template <int HD>
static int launch_bwd(const float* dO, int lddo, const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv,
                      const float* A, float* dA, int ldA, float* delta, float* dQ, int lddq, float* dK, int lddk, float* dV,
                      int lddv, int B, int H, int T, int S, float scale, cudaStream_t st) {
  const size_t smem = AttnSmem<HD>::bytes(ldA);
  MMX_REQUIRE(smem <= 227 * 1024, "sequence too long for the single-pass attention kernel");
  MMX_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_q_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(T, TQ), H, B);
  attention_bwd_q_kernel<HD><<<grid, ATT_THREADS, smem, st>>>(dO, lddo, K, ldk, V, ldv, A, dA, ldA, delta, dQ, lddq, H, T, S,
                                                              scale);
  MMX_LAUNCH_CHECK();
  if (dQ == nullptr) return 0;
  const size_t smem2 = sizeof(float) * (2 * KV_ROWS * (HD + 1) + 2 * KV_ROWS * (KV_KEYS + 1));
  MMX_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  dim3 grid2(cdiv(S, KV_KEYS), H, B);
  attention_bwd_kv_kernel<HD><<<grid2, ATT_THREADS, smem2, st>>>(dO, lddo, Q, ldq, A, dA, ldA, delta, dK, lddk, dV, lddv, H,
                                                                 T, S, scale);
  MMX_LAUNCH_CHECK();
  return 0;
}

int attention_fwd(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const float* key_bias, float* A,
                  int ldA, float* O, int ldo, int B, int H, int T, int S, int hd, float scale, int flags, cudaStream_t st) {
  MMX_REQUIRE(ldA >= S, "ldA < S");
  if (B == 0 || T == 0) return 0;
  switch (hd) {
    case 16: return launch_fwd<16>(Q, ldq, K, ldk, V, ldv, key_bias, A, ldA, O, ldo, B, H, T, S, scale, flags, st);
    case 32: return launch_fwd<32>(Q, ldq, K, ldk, V, ldv, key_bias, A, ldA, O, ldo, B, H, T, S, scale, flags, st);
    case 64: return launch_fwd<64>(Q, ldq, K, ldk, V, ldv, key_bias, A, ldA, O, ldo, B, H, T, S, scale, flags, st);
    default: MMX_REQUIRE(false, "head_dim must be 16, 32 or 64");
  }
  return 0;
}

int attention_bwd(const float* dO, int lddo, const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv,
                  const float* A, float* dA, int ldA, float* delta, float* dQ, int lddq, float* dK, int lddk, float* dV,
                  int lddv, int B, int H, int T, int S, int hd, float scale, int flags, cudaStream_t st) {
  MMX_REQUIRE(ldA >= S, "ldA < S");
  MMX_REQUIRE((dQ == nullptr) == (dK == nullptr) && (dQ == nullptr) == (dV == nullptr), "dQ/dK/dV: all or none");
  MMX_REQUIRE(dQ == nullptr || delta != nullptr, "delta scratch required");
  (void)flags;  // the mask is already folded into A (masked entries are exactly 0, so dS vanishes there)
  if (B == 0 || T == 0) return 0;
  switch (hd) {
    case 16: return launch_bwd<16>(dO, lddo, Q, ldq, K, ldk, V, ldv, A, dA, ldA, delta, dQ, lddq, dK, lddk, dV, lddv, B, H, T, S, scale, st);
    case 32: return launch_bwd<32>(dO, lddo, Q, ldq, K, ldk, V, ldv, A, dA, ldA, delta, dQ, lddq, dK, lddk, dV, lddv, B, H, T, S, scale, st);
    case 64: return launch_bwd<64>(dO, lddo, Q, ldq, K, ldk, V, ldv, A, dA, ldA, delta, dQ, lddq, dK, lddk, dV, lddv, B, H, T, S, scale, st);
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
  return attention_fwd(Q, ldq, K, ldk, V, ldv, key_bias, A, ldA, O, ldo, B, H, T, S, hd, scale, flags, (cudaStream_t)stream);
}
int mmx_attention_bwd(const float* dO, int lddo, const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv,
                      const float* A, float* dA, int ldA, float* delta, float* dQ, int lddq, float* dK, int lddk, float* dV,
                      int lddv, int B, int H, int T, int S, int hd, float scale, int flags, void* stream) {
  return attention_bwd(dO, lddo, Q, ldq, K, ldk, V, ldv, A, dA, ldA, delta, dQ, lddq, dK, lddk, dV, lddv, B, H, T, S, hd, scale,
                       flags, (cudaStream_t)stream);
}
}
