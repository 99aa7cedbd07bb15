// Coalesced epilogue of the tcgen05 GEMM kernels.
//
// tcgen05.ld hands every thread ONE ROW of the accumulator tile (TMEM lane = row).  Storing from that layout directly
// (thread -> 16 B of its own row) makes each warp-wide store touch 32 different rows, 16 B each: half-filled 32-byte
// sectors, 32 requests per instruction.  Measured on the QKV product (3200 x 2304 x 768, clock64 timeline of CTA 0,
// profiles/gemm_trace_r2_before.log): 19300 clk of stores per 128 x 144 tile against 13000 clk of main loop - the epilogue, not
// the tensor pipe, L2 or the splitters, set the time of every variant of the kernel.
//
// Here each epilogue warp turns its 32 rows x 16 columns through a 2 KB shared-memory buffer (XOR-swizzled 16-byte chunks:
// conflict-free both ways) so that a warp instruction covers 8 rows x 64 contiguous bytes - whole sectors - for the C
// store, the act(C) store and the residual / pre-activation loads alike.
#pragma once
#include "gemm.cuh"
#include "mmx_common.cuh"

namespace mmx {

constexpr int EPI_STAGE_FLOATS = 32 * 16;      // per epilogue warp

// One copy of the activation bodies per kernel (called, not inlined at every use: see the note on code size below).
static __device__ __noinline__ float4 act_bwd4(float4 f, int act) {
  return make_float4(act_bwd(f.x, act), act_bwd(f.y, act), act_bwd(f.z, act), act_bwd(f.w, act));
}
static __device__ __noinline__ float4 act_fwd4(float4 v, int act) {
  return make_float4(act_fwd(v.x, act), act_fwd(v.y, act), act_fwd(v.z, act), act_fwd(v.w, act));
}

// acc: this thread's row (m_warp0 + lane) of the warp's 32 x CW block, columns nbase .. nbase + CW - 1 (CW % 8 == 0).
// stg: the warp's private staging buffer (EPI_STAGE_FLOATS floats, 16-byte aligned).
// ACT = false: bias / residual only - no activation code in the kernel at all.  The activation variants keep the row loop
// rolled: the erf / exp / tanh bodies are large, and an epilogue that outgrows the instruction cache stalls on every
// fetch (the first coalesced version, fully unrolled with the activation switch inlined 80 times, was a 64 KB kernel and
// still spent 10-16 k clk per tile in this function).
template <int CW, bool ACT>
__device__ __forceinline__ void epilogue_store_rows(const uint32_t (&acc)[CW], float* stg, int m_warp0, int nbase, int M, int N,
                                                    float* C, int ldc, long long c_off, long long res_off, const GemmEpilogue& ep,
                                                    int lane) {
  static_assert(CW % 8 == 0, "column block");
  const int sw_w = (lane >> 1) & 3;                 // swizzle of the row this thread WRITES (row = lane)
  const int rr = lane >> 2, cq = lane & 3;          // read side: row rr + 8 i, 16-byte chunk cq
#pragma unroll
  for (int c0 = 0; c0 < CW; c0 += 16) {
    constexpr int FULL = 4;
    const int chunks = (CW - c0 >= 16) ? FULL : (CW - c0) / 4;       // 4, or 2 in the 8-wide tail of CW = 72
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < chunks) {
        const float4 v = make_float4(__uint_as_float(acc[c0 + 4 * j]), __uint_as_float(acc[c0 + 4 * j + 1]),
                                     __uint_as_float(acc[c0 + 4 * j + 2]), __uint_as_float(acc[c0 + 4 * j + 3]));
        *reinterpret_cast<float4*>(stg + lane * 16 + ((j ^ sw_w) << 2)) = v;
      }
    }
    __syncwarp();
    const int n = nbase + c0 + 4 * cq;
    if (cq < chunks && n < N) {                                       // N % 4 == 0: a chunk is inside or outside
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ep.bias) b = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
#pragma unroll ACT ? 1 : 4
      for (int i = 0; i < 4; ++i) {
        const int r = rr + 8 * i, m = m_warp0 + r;
        if (m < M) {
          float4 v = *reinterpret_cast<const float4*>(stg + r * 16 + ((cq ^ ((r >> 1) & 3)) << 2));
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
          if (ACT && ep.pre) {
            const float4 f = act_bwd4(*reinterpret_cast<const float4*>(ep.pre + (long long)m * ep.ldpre + n), ep.act);
            v.x *= f.x; v.y *= f.y; v.z *= f.z; v.w *= f.w;
          }
          if (ep.residual) {
            const float4 a = *reinterpret_cast<const float4*>(ep.residual + res_off + (long long)m * ep.ldres + n);
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
          }
          *reinterpret_cast<float4*>(C + c_off + (long long)m * ldc + n) = v;
          if (ACT && ep.C_act) {
            *reinterpret_cast<float4*>(ep.C_act + (long long)m * ldc + n) = act_fwd4(v, ep.act);
          }
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace mmx
