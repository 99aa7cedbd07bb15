// Layer-wise relevance propagation (the `relprop` sweep behind use_lrp=True) as device kernels.
//
// The reference implements relprop as gradient-x-input autograd calls per layer (DETR/modules/layers.py:38-66,
// lxmert/lxmert/src/layers.py, VisualBERT/.../layers_ours.py).  Here every rule is a direct kernel:
//   Linear  (alpha = 1, beta = 0):  Z = x+ W+^T + x- W-^T   (two GEMMs, gemm_dispatch.cu)
//                                   S = safe_divide(R, Z)    (safe_divide_kernel)
//                                   out = x+ (.) (S W+) + x- (.) (S W-)   (two GEMMs whose epilogue multiplies by `pre`, MMX_ACT_MUL)
//                                   DETR only: out *= safe_divide(sum R, sum out) per sample  (sums_kernel + renorm_kernel)
//   Add / Clone / IndexSelect:      add_pass1/2, clone_kernel
//   attention (two RelPropSimple matmuls, layers.py:770-801 / lxmert_lrp.py:422-461): attn_pv_*, attn_qk_* below; the
//   relevance of the attention probabilities (`attn_cam`) is staged in the [B,H,T,ld] layout rule 5 reads.
// Sums over a whole sample are deterministic: 64 per-block partials in double, reduced in a fixed order by the consumer.
// One "sample" = the reference's whole tensor (it runs relprop with batch 1); a batch is B independent samples.
#include "mmx_common.cuh"

namespace mmx {
namespace lrp {

constexpr int NCH = 64;   // partial sums per sample

// layers.py:11-14  den = b.clamp(min=1e-9) + b.clamp(max=1e-9); den += (den == 0) * 1e-9; a / den * (b != 0)
__device__ __forceinline__ float safe_div(float a, float b) {
  float den = (b != b) ? b : (fmaxf(b, 1e-9f) + fminf(b, 1e-9f));
  if (den == 0.f) den += 1e-9f;
  return (a / den) * (b != 0.f ? 1.f : 0.f);
}

__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ X, int ldx, float* __restrict__ P,
                                                    float* __restrict__ Ng, int ld, long long rows, int cols) {
  const long long n = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    const float x = X[r * ldx + c];
    P[r * ld + c] = (x != x) ? x : fmaxf(x, 0.f);     // clamp(min=0) / clamp(max=0) propagate NaN
    Ng[r * ld + c] = (x != x) ? x : fminf(x, 0.f);
  }
}

__global__ void __launch_bounds__(256) safe_divide_kernel(const float* __restrict__ R, int ldr, const float* __restrict__ Z, int ldz,
                                                          float* __restrict__ S, int lds, long long rows, int cols) {
  const long long n = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    S[r * lds + c] = safe_div(R[r * ldr + c], Z[r * ldz + c]);
  }
}

// block-wide sum of NV doubles per thread -> out[0..NV) valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* out) {
  __shared__ double sh[NV][8];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) sh[k][warp] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[k][w];
      out[k] = s;
    }
  }
}

// element range of chunk c of a sample with n elements
__device__ __forceinline__ void chunk_range(long long n, int c, long long& lo, long long& hi) {
  const long long per = (n + NCH - 1) / NCH;
  lo = per * c;
  hi = lo + per < n ? lo + per : n;
  if (lo > n) lo = n;
}

// part[b][c] = {sum, abs-sum} of chunk c of sample b
__global__ void __launch_bounds__(256) sums_kernel(const float* __restrict__ X, int ldx, int rows_ps, int cols,
                                                   double* __restrict__ part) {
  const int b = blockIdx.y, c = blockIdx.x;
  const long long n = (long long)rows_ps * cols;
  long long lo, hi;
  chunk_range(n, c, lo, hi);
  const float* Xb = X + (long long)b * rows_ps * ldx;
  double v[2] = {0, 0};
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const long long r = i / cols;
    const float x = Xb[r * ldx + (int)(i - r * cols)];
    v[0] += (double)x;
    v[1] += (double)fabsf(x);
  }
  block_sum<2>(v, part + ((long long)b * NCH + c) * 2);
}

__device__ __forceinline__ double reduce_part(const double* part, int b, int stride, int k) {
  double s = 0;
  for (int c = 0; c < NCH; ++c) s += part[((long long)b * NCH + c) * stride + k];
  return s;
}

// X[b] *= safe_divide(sum(num[b]), sum(den[b]))   (layers.py:430-432: `out * safe_divide(R.sum(), out.sum())`)
__global__ void __launch_bounds__(256) renorm_kernel(float* __restrict__ X, int ldx, int rows_ps, int cols,
                                                     const double* __restrict__ num, const double* __restrict__ den) {
  const int b = blockIdx.y;
  __shared__ float f;
  if (threadIdx.x == 0) f = safe_div((float)reduce_part(num, b, 2, 0), (float)reduce_part(den, b, 2, 0));
  __syncthreads();
  const float fac = f;
  const long long n = (long long)rows_ps * cols;
  float* Xb = X + (long long)b * rows_ps * ldx;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    Xb[r * ldx + (int)(i - r * cols)] *= fac;
  }
}

// Add.relprop (layers.py:194-221), pass 1: a = x0 * S, b = x1 * S with S = safe_divide(R, x0 + x1); partial sums of a, b, R
__global__ void __launch_bounds__(256) add_pass1_kernel(const float* __restrict__ R, int ldr, const float* __restrict__ x0, int ld0,
                                                        const float* __restrict__ x1, int ld1, float* __restrict__ a, int lda,
                                                        float* __restrict__ bo, int ldb, int rows_ps, int cols,
                                                        double* __restrict__ part) {
  const int b = blockIdx.y, c = blockIdx.x;
  const long long n = (long long)rows_ps * cols;
  long long lo, hi;
  chunk_range(n, c, lo, hi);
  const long long r0 = (long long)b * rows_ps;
  double v[3] = {0, 0, 0};
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const long long r = r0 + i / cols;
    const int cc = (int)(i % cols);
    const float rr = R[r * ldr + cc], u0 = x0[r * ld0 + cc], u1 = x1[r * ld1 + cc];
    const float s = safe_div(rr, u0 + u1);
    const float av = u0 * s, bv = u1 * s;
    a[r * lda + cc] = av;
    if (bo) bo[r * ldb + cc] = bv;
    v[0] += (double)av; v[1] += (double)bv; v[2] += (double)rr;
  }
  block_sum<3>(v, part + ((long long)b * NCH + c) * 3);
}
// pass 2: a *= safe_divide(a_fact, a_sum), a_fact = safe_divide(|a_sum|, |a_sum| + |b_sum|) * R_sum (same for b)
__global__ void __launch_bounds__(256) add_pass2_kernel(float* __restrict__ a, int lda, float* __restrict__ bo, int ldb, int rows_ps,
                                                        int cols, const double* __restrict__ part) {
  const int b = blockIdx.y;
  __shared__ float fa, fb;
  if (threadIdx.x == 0) {
    const float as = (float)reduce_part(part, b, 3, 0), bs = (float)reduce_part(part, b, 3, 1), rs = (float)reduce_part(part, b, 3, 2);
    const float a_fact = safe_div(fabsf(as), fabsf(as) + fabsf(bs)) * rs;
    const float b_fact = safe_div(fabsf(bs), fabsf(as) + fabsf(bs)) * rs;
    fa = safe_div(a_fact, as);
    fb = safe_div(b_fact, bs);
  }
  __syncthreads();
  const float ka = fa, kb = fb;
  const long long n = (long long)rows_ps * cols, r0 = (long long)b * rows_ps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = r0 + i / cols;
    const int cc = (int)(i % cols);
    a[r * lda + cc] *= ka;
    if (bo) bo[r * ldb + cc] *= kb;
  }
}

// Clone.relprop (layers.py:252-270): out = X * sum_i safe_divide(R_i, X)
struct ClonePtrs { const float* r[8]; int ld[8]; };
__global__ void __launch_bounds__(256) clone_kernel(const float* __restrict__ X, int ldx, ClonePtrs p, int n_in, float* __restrict__ out,
                                                    int ldo, long long rows, int cols) {
  const long long n = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    const float x = X[r * ldx + c];
    float s = 0.f;
    for (int k = 0; k < n_in; ++k) s += safe_div(p.r[k][r * p.ld[k] + c], x);
    out[r * ldo + c] = x * s;
  }
}

// The all-zero value special case of MultiheadAttention.relprop (layers.py:790-799): when the value branch received
// relevance (its relevance was not all zero before the value projection) but the projection's input is all zero, the
// query / key relevances are rescaled to share the relevance `cam` that entered the attention.
__global__ void __launch_bounds__(256) zero_value_fix_kernel(float* __restrict__ cq, int ldq, int rows_q, float* __restrict__ ck, int ldk,
                                                             int rows_k, int cols, const double* __restrict__ pv_pre,
                                                             const double* __restrict__ pv_post, const double* __restrict__ pq,
                                                             const double* __restrict__ pk, const double* __restrict__ pcam) {
  const int b = blockIdx.y;
  __shared__ float fq, fk;
  __shared__ int apply;
  if (threadIdx.x == 0) {
    const bool pre_zero = reduce_part(pv_pre, b, 2, 1) == 0.0, post_zero = reduce_part(pv_post, b, 2, 1) == 0.0;
    apply = post_zero && !pre_zero;
    const float ks = (float)reduce_part(pk, b, 2, 0), qs = (float)reduce_part(pq, b, 2, 0), cs = (float)reduce_part(pcam, b, 2, 0);
    const float k_fact = safe_div(fabsf(ks), fabsf(ks) + fabsf(qs)) * cs;
    const float q_fact = safe_div(fabsf(qs), fabsf(ks) + fabsf(qs)) * cs;
    fk = safe_div(k_fact, ks);
    fq = safe_div(q_fact, qs);
  }
  __syncthreads();
  if (!apply) return;
  const long long nq = (long long)rows_q * cols, nk = (long long)rows_k * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (long long)gridDim.x * blockDim.x) {
    const long long r = (long long)b * rows_q + i / cols;
    cq[r * ldq + (int)(i % cols)] *= fq;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nk; i += (long long)gridDim.x * blockDim.x) {
    const long long r = (long long)b * rows_k + i / cols;
    ck[r * ldk + (int)(i % cols)] *= fk;
  }
}

// ------------------------------------------------------------------------------------------------ attention relprop
// Head h of sample b: q/k/v/o rows b*T+i (or b*S+j), columns h*hd .. h*hd+hd-1; A and the staged relevances [B,H,T,ldA].
constexpr int TQ = 32, TK = 64, HDMAX = 64;

// relevance of the probabilities: cam_A[i,j] = A[i,j] * sum_d Sv[i,d] v[j,d] / 2,  Sv = safe_divide(R_o, o)
__global__ void __launch_bounds__(256) attn_pv_a_kernel(const float* __restrict__ Ro, int ldr, const float* __restrict__ O, int ldo,
                                                        const float* __restrict__ A, const float* __restrict__ V, int ldv,
                                                        float* __restrict__ camA, int ldA, int H, int T, int S, int hd) {
  __shared__ float sS[TQ][HDMAX + 1];
  __shared__ float sV[TK][HDMAX + 1];
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * TQ;
  const int tid = threadIdx.x;
  for (int e = tid; e < TQ * hd; e += 256) {
    const int i = e / hd, d = e % hd;
    float s = 0.f;
    if (i0 + i < T) {
      const long long row = (long long)b * T + i0 + i;
      s = safe_div(Ro[row * ldr + h * hd + d], O[row * ldo + h * hd + d]);
    }
    sS[i][d] = s;
  }
  const float* Ah = A + ((long long)(b * H + h) * T) * ldA;
  float* Ch = camA + ((long long)(b * H + h) * T) * ldA;
  for (int j0 = 0; j0 < ldA; j0 += TK) {
    __syncthreads();
    for (int e = tid; e < TK * hd; e += 256) {
      const int j = e / hd, d = e % hd;
      sV[j][d] = (j0 + j < S) ? V[((long long)b * S + j0 + j) * ldv + h * hd + d] : 0.f;
    }
    __syncthreads();
    const int j = tid & 63;
    for (int i = tid >> 6; i < TQ; i += 4) {
      if (i0 + i >= T || j0 + j >= ldA) continue;
      float out = 0.f;
      if (j0 + j < S) {
        float acc = 0.f;
        for (int d = 0; d < hd; ++d) acc = fmaf(sS[i][d], sV[j][d], acc);
        out = Ah[(long long)(i0 + i) * ldA + j0 + j] * acc * 0.5f;
      }
      Ch[(long long)(i0 + i) * ldA + j0 + j] = out;       // pad columns S..ldA-1 are written as zeros
    }
  }
}

// relevance of the values: cam_v[j,d] = v[j,d] * sum_i A[i,j] Sv[i,d] / 2
__global__ void __launch_bounds__(256) attn_pv_v_kernel(const float* __restrict__ Ro, int ldr, const float* __restrict__ O, int ldo,
                                                        const float* __restrict__ A, int ldA, const float* __restrict__ V, int ldv,
                                                        float* __restrict__ camV, int ldcv, int H, int T, int S, int hd) {
  __shared__ float sS[TQ][HDMAX + 1];
  __shared__ float sA[TQ][TK + 1];
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * TK;
  const int tid = threadIdx.x;
  const float* Ah = A + ((long long)(b * H + h) * T) * ldA;
  // thread -> (key j, a strided set of d): j = tid & 63, d = (tid >> 6) + 4 * m
  const int j = tid & 63, dq = tid >> 6;
  float acc[HDMAX / 4];
#pragma unroll
  for (int m = 0; m < HDMAX / 4; ++m) acc[m] = 0.f;
  for (int i0 = 0; i0 < T; i0 += TQ) {
    __syncthreads();
    for (int e = tid; e < TQ * hd; e += 256) {
      const int i = e / hd, d = e % hd;
      float s = 0.f;
      if (i0 + i < T) {
        const long long row = (long long)b * T + i0 + i;
        s = safe_div(Ro[row * ldr + h * hd + d], O[row * ldo + h * hd + d]);
      }
      sS[i][d] = s;
    }
    for (int e = tid; e < TQ * TK; e += 256) {
      const int i = e / TK, jj = e % TK;
      sA[i][jj] = (i0 + i < T && j0 + jj < S) ? Ah[(long long)(i0 + i) * ldA + j0 + jj] : 0.f;
    }
    __syncthreads();
    for (int i = 0; i < TQ; ++i) {
      const float a = sA[i][j];
#pragma unroll
      for (int m = 0; m < HDMAX / 4; ++m) {
        const int d = dq + 4 * m;
        if (d < hd) acc[m] = fmaf(a, sS[i][d], acc[m]);
      }
    }
  }
  if (j0 + j < S) {
    const long long row = (long long)b * S + j0 + j;
#pragma unroll
    for (int m = 0; m < HDMAX / 4; ++m) {
      const int d = dq + 4 * m;
      if (d < hd) camV[row * ldcv + h * hd + d] = V[row * ldv + h * hd + d] * acc[m] * 0.5f;
    }
  }
}

// relevance of the queries: cam_q[i,d] = (zs q[i,d]) * sum_j S2[i,j] k[j,d] / 2,  S2 = safe_divide(cam1, zs * q k^T)
__global__ void __launch_bounds__(256) attn_qk_q_kernel(const float* __restrict__ cam1, int ldA, const float* __restrict__ Q, int ldq,
                                                        const float* __restrict__ K, int ldk, float zs, float* __restrict__ camQ,
                                                        int ldcq, int H, int T, int S, int hd) {
  __shared__ float sQ[TQ][HDMAX + 1];
  __shared__ float sK[TK][HDMAX + 1];
  __shared__ float sS2[TQ][TK + 1];
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * TQ;
  const int tid = threadIdx.x;
  for (int e = tid; e < TQ * hd; e += 256) {
    const int i = e / hd, d = e % hd;
    sQ[i][d] = (i0 + i < T) ? zs * Q[((long long)b * T + i0 + i) * ldq + h * hd + d] : 0.f;
  }
  const float* Ch = cam1 + ((long long)(b * H + h) * T) * ldA;
  // output mapping: i = tid >> 3 (32 rows), d = (tid & 7) + 8 * m
  const int oi = tid >> 3, od = tid & 7;
  float acc[HDMAX / 8];
#pragma unroll
  for (int m = 0; m < HDMAX / 8; ++m) acc[m] = 0.f;
  for (int j0 = 0; j0 < S; j0 += TK) {
    __syncthreads();
    for (int e = tid; e < TK * hd; e += 256) {
      const int j = e / hd, d = e % hd;
      sK[j][d] = (j0 + j < S) ? K[((long long)b * S + j0 + j) * ldk + h * hd + d] : 0.f;
    }
    __syncthreads();
    {
      const int j = tid & 63;
      for (int i = tid >> 6; i < TQ; i += 4) {
        float s2 = 0.f;
        if (i0 + i < T && j0 + j < S) {
          float z = 0.f;
          for (int d = 0; d < hd; ++d) z = fmaf(sQ[i][d], sK[j][d], z);
          s2 = safe_div(Ch[(long long)(i0 + i) * ldA + j0 + j], z);
        }
        sS2[i][j] = s2;
      }
    }
    __syncthreads();
    for (int j = 0; j < TK; ++j) {
      const float s2 = sS2[oi][j];
#pragma unroll
      for (int m = 0; m < HDMAX / 8; ++m) {
        const int d = od + 8 * m;
        if (d < hd) acc[m] = fmaf(s2, sK[j][d], acc[m]);
      }
    }
  }
  if (i0 + oi < T) {
    const long long row = (long long)b * T + i0 + oi;
#pragma unroll
    for (int m = 0; m < HDMAX / 8; ++m) {
      const int d = od + 8 * m;
      if (d < hd) camQ[row * ldcq + h * hd + d] = sQ[oi][d] * acc[m] * 0.5f;
    }
  }
}

// relevance of the keys: cam_k[j,d] = k[j,d] * sum_i S2[i,j] (zs q[i,d]) / 2
__global__ void __launch_bounds__(256) attn_qk_k_kernel(const float* __restrict__ cam1, int ldA, const float* __restrict__ Q, int ldq,
                                                        const float* __restrict__ K, int ldk, float zs, float* __restrict__ camK,
                                                        int ldck, int H, int T, int S, int hd) {
  __shared__ float sQ[TQ][HDMAX + 1];
  __shared__ float sK[TK][HDMAX + 1];
  __shared__ float sS2[TQ][TK + 1];
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * TK;
  const int tid = threadIdx.x;
  for (int e = tid; e < TK * hd; e += 256) {
    const int j = e / hd, d = e % hd;
    sK[j][d] = (j0 + j < S) ? K[((long long)b * S + j0 + j) * ldk + h * hd + d] : 0.f;
  }
  const float* Ch = cam1 + ((long long)(b * H + h) * T) * ldA;
  const int j = tid & 63, dq = tid >> 6;
  float acc[HDMAX / 4];
#pragma unroll
  for (int m = 0; m < HDMAX / 4; ++m) acc[m] = 0.f;
  for (int i0 = 0; i0 < T; i0 += TQ) {
    __syncthreads();
    for (int e = tid; e < TQ * hd; e += 256) {
      const int i = e / hd, d = e % hd;
      sQ[i][d] = (i0 + i < T) ? zs * Q[((long long)b * T + i0 + i) * ldq + h * hd + d] : 0.f;
    }
    __syncthreads();
    for (int i = tid >> 6; i < TQ; i += 4) {
      float s2 = 0.f;
      if (i0 + i < T && j0 + j < S) {
        float z = 0.f;
        for (int d = 0; d < hd; ++d) z = fmaf(sQ[i][d], sK[j][d], z);
        s2 = safe_div(Ch[(long long)(i0 + i) * ldA + j0 + j], z);
      }
      sS2[i][j] = s2;
    }
    __syncthreads();
    for (int i = 0; i < TQ; ++i) {
      const float s2 = sS2[i][j];
#pragma unroll
      for (int m = 0; m < HDMAX / 4; ++m) {
        const int d = dq + 4 * m;
        if (d < hd) acc[m] = fmaf(s2, sQ[i][d], acc[m]);
      }
    }
  }
  if (j0 + j < S) {
    const long long row = (long long)b * S + j0 + j;
#pragma unroll
    for (int m = 0; m < HDMAX / 4; ++m) {
      const int d = dq + 4 * m;
      if (d < hd) camK[row * ldck + h * hd + d] = sK[j][d] * acc[m] * 0.5f;
    }
  }
}

// scores[b,h,i,j] = zs * q.k (+ bias[b,j]) written in the [B,H,T,ld] layout; mask[b,h,i,j] = bias[b,j] (the additive
// attention mask expanded as the reference's Add sees it).  VisualBERT's BertSelfAttention.relprop sends the relevance of
// the probabilities through that Add (BERT_ours.py:352-395).
__global__ void __launch_bounds__(256) attn_scores_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ K, int ldk,
                                                          const float* __restrict__ bias, float zs, float* __restrict__ scores,
                                                          float* __restrict__ mask, int ldA, int H, int T, int S, int hd) {
  __shared__ float sQ[TQ][HDMAX + 1];
  __shared__ float sK[TK][HDMAX + 1];
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * TQ;
  const int tid = threadIdx.x;
  for (int e = tid; e < TQ * hd; e += 256) {
    const int i = e / hd, d = e % hd;
    sQ[i][d] = (i0 + i < T) ? Q[((long long)b * T + i0 + i) * ldq + h * hd + d] : 0.f;
  }
  float* Sh = scores + ((long long)(b * H + h) * T) * ldA;
  float* Mh = mask ? mask + ((long long)(b * H + h) * T) * ldA : nullptr;
  for (int j0 = 0; j0 < ldA; j0 += TK) {
    __syncthreads();
    for (int e = tid; e < TK * hd; e += 256) {
      const int j = e / hd, d = e % hd;
      sK[j][d] = (j0 + j < S) ? K[((long long)b * S + j0 + j) * ldk + h * hd + d] : 0.f;
    }
    __syncthreads();
    const int j = tid & 63;
    for (int i = tid >> 6; i < TQ; i += 4) {
      if (i0 + i >= T || j0 + j >= ldA) continue;
      float z = 0.f, mb = 0.f;
      if (j0 + j < S) {
        for (int d = 0; d < hd; ++d) z = fmaf(sQ[i][d], sK[j][d], z);
        z *= zs;
        if (bias) mb = bias[(long long)b * S + j0 + j];
      }
      Sh[(long long)(i0 + i) * ldA + j0 + j] = z;
      if (Mh) Mh[(long long)(i0 + i) * ldA + j0 + j] = mb;
    }
  }
}

inline int grid_for(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace lrp
}  // namespace mmx

using namespace mmx;
using namespace mmx::lrp;

extern "C" {

int mmx_lrp_split(const float* X, int ldx, float* P, float* Ng, int ld, long long rows, int cols, void* stream) {
  if (rows == 0 || cols == 0) return 0;
  MMX_REQUIRE(X && P && Ng && ldx >= cols && ld >= cols, "bad arguments");
  split_kernel<<<grid_for(rows * cols), 256, 0, (cudaStream_t)stream>>>(X, ldx, P, Ng, ld, rows, cols);
  MMX_LAUNCH_CHECK();
  return 0;
}

int mmx_lrp_safe_divide(const float* R, int ldr, const float* Z, int ldz, float* S, int lds, long long rows, int cols, void* stream) {
  if (rows == 0 || cols == 0) return 0;
  MMX_REQUIRE(R && Z && S, "bad arguments");
  safe_divide_kernel<<<grid_for(rows * cols), 256, 0, (cudaStream_t)stream>>>(R, ldr, Z, ldz, S, lds, rows, cols);
  MMX_LAUNCH_CHECK();
  return 0;
}

int mmx_lrp_partials_len(void) { return NCH; }

int mmx_lrp_sums(const float* X, int ldx, int rows_per_sample, int cols, int B, double* partials, void* stream) {
  if (B == 0) return 0;
  MMX_REQUIRE(X && partials && rows_per_sample >= 0 && cols >= 0, "bad arguments");
  sums_kernel<<<dim3(NCH, B), 256, 0, (cudaStream_t)stream>>>(X, ldx, rows_per_sample, cols, partials);
  MMX_LAUNCH_CHECK();
  return 0;
}

int mmx_lrp_renorm(float* X, int ldx, int rows_per_sample, int cols, int B, const double* num_partials, const double* den_partials,
                   void* stream) {
  if (B == 0 || rows_per_sample == 0 || cols == 0) return 0;
  MMX_REQUIRE(X && num_partials && den_partials, "bad arguments");
  renorm_kernel<<<dim3(grid_for((long long)rows_per_sample * cols), B), 256, 0, (cudaStream_t)stream>>>(X, ldx, rows_per_sample, cols,
                                                                                                      num_partials, den_partials);
  MMX_LAUNCH_CHECK();
  return 0;
}

int mmx_lrp_add(const float* R, int ldr, const float* x0, int ld0, const float* x1, int ld1, float* a, int lda, float* b, int ldb,
                int rows_per_sample, int cols, int B, double* workspace, void* stream) {
  if (B == 0 || rows_per_sample == 0 || cols == 0) return 0;
  MMX_REQUIRE(R && x0 && x1 && a && workspace, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  add_pass1_kernel<<<dim3(NCH, B), 256, 0, st>>>(R, ldr, x0, ld0, x1, ld1, a, lda, b, ldb, rows_per_sample, cols, workspace);
  MMX_LAUNCH_CHECK();
  add_pass2_kernel<<<dim3(grid_for((long long)rows_per_sample * cols), B), 256, 0, st>>>(a, lda, b, ldb, rows_per_sample, cols, workspace);
  MMX_LAUNCH_CHECK();
  return 0;
}

int mmx_lrp_clone(const float* X, int ldx, const float* const* R, const int* ldr, int n, float* out, int ldo, long long rows, int cols,
                  void* stream) {
  if (rows == 0 || cols == 0) return 0;
  MMX_REQUIRE(X && R && ldr && out && n >= 1 && n <= 8, "1 to 8 relevance inputs");
  ClonePtrs p;
  for (int i = 0; i < 8; ++i) { p.r[i] = i < n ? R[i] : nullptr; p.ld[i] = i < n ? ldr[i] : 0; }
  for (int i = 0; i < n; ++i) MMX_REQUIRE(p.r[i] != nullptr, "null relevance input");
  clone_kernel<<<grid_for(rows * cols), 256, 0, (cudaStream_t)stream>>>(X, ldx, p, n, out, ldo, rows, cols);
  MMX_LAUNCH_CHECK();
  return 0;
}

int mmx_lrp_zero_value_fix(float* cam_q, int ldq, int rows_q, float* cam_k, int ldk, int rows_k, int cols, int B,
                           const double* v_pre, const double* v_post, const double* q_sums, const double* k_sums,
                           const double* cam_sums, void* stream) {
  if (B == 0) return 0;
  MMX_REQUIRE(cam_q && cam_k && v_pre && v_post && q_sums && k_sums && cam_sums, "bad arguments");
  const long long n = (long long)(rows_q > rows_k ? rows_q : rows_k) * cols;
  zero_value_fix_kernel<<<dim3(grid_for(n), B), 256, 0, (cudaStream_t)stream>>>(cam_q, ldq, rows_q, cam_k, ldk, rows_k, cols, v_pre,
                                                                               v_post, q_sums, k_sums, cam_sums);
  MMX_LAUNCH_CHECK();
  return 0;
}

int mmx_lrp_attn_pv(const float* R_o, int ldr, const float* O, int ldo, const float* A, int ldA, const float* V, int ldv,
                    float* cam_A, float* cam_V, int ldcv, int B, int H, int T, int S, int hd, void* stream) {
  if (B == 0 || T == 0 || S == 0) return 0;
  MMX_REQUIRE(R_o && O && A && V && cam_A && cam_V, "bad arguments");
  MMX_REQUIRE(hd >= 1 && hd <= HDMAX && ldA >= S, "head_dim <= 64");
  cudaStream_t st = (cudaStream_t)stream;
  attn_pv_a_kernel<<<dim3(cdiv(T, TQ), H, B), 256, 0, st>>>(R_o, ldr, O, ldo, A, V, ldv, cam_A, ldA, H, T, S, hd);
  MMX_LAUNCH_CHECK();
  attn_pv_v_kernel<<<dim3(cdiv(S, TK), H, B), 256, 0, st>>>(R_o, ldr, O, ldo, A, ldA, V, ldv, cam_V, ldcv, H, T, S, hd);
  MMX_LAUNCH_CHECK();
  return 0;
}

int mmx_lrp_attn_qk(const float* cam1, int ldA, const float* Q, int ldq, const float* K, int ldk, float zscale, float* cam_Q,
                    int ldcq, float* cam_K, int ldck, int B, int H, int T, int S, int hd, void* stream) {
  if (B == 0 || T == 0 || S == 0) return 0;
  MMX_REQUIRE(cam1 && Q && K && cam_Q && cam_K, "bad arguments");
  MMX_REQUIRE(hd >= 1 && hd <= HDMAX && ldA >= S, "head_dim <= 64");
  cudaStream_t st = (cudaStream_t)stream;
  attn_qk_q_kernel<<<dim3(cdiv(T, TQ), H, B), 256, 0, st>>>(cam1, ldA, Q, ldq, K, ldk, zscale, cam_Q, ldcq, H, T, S, hd);
  MMX_LAUNCH_CHECK();
  attn_qk_k_kernel<<<dim3(cdiv(S, TK), H, B), 256, 0, st>>>(cam1, ldA, Q, ldq, K, ldk, zscale, cam_K, ldck, H, T, S, hd);
  MMX_LAUNCH_CHECK();
  return 0;
}

int mmx_lrp_attn_scores(const float* Q, int ldq, const float* K, int ldk, const float* key_bias, float zscale, float* scores,
                        float* mask, int ldA, int B, int H, int T, int S, int hd, void* stream) {
  if (B == 0 || T == 0 || S == 0) return 0;
  MMX_REQUIRE(Q && K && scores, "bad arguments");
  MMX_REQUIRE(hd >= 1 && hd <= HDMAX && ldA >= S, "head_dim <= 64");
  attn_scores_kernel<<<dim3(cdiv(T, TQ), H, B), 256, 0, (cudaStream_t)stream>>>(Q, ldq, K, ldk, key_bias, zscale, scores, mask, ldA, H,
                                                                               T, S, hd);
  MMX_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
