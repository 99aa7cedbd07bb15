// Rule 6 as ONE launch per tower:  R = I;  for l in layers:  R <- R + Abar_l R     (CLIP_explainability.ipynb:169-183,
// Transformer_MM_explainability_ViT.ipynb:1193-1199, VisualBERT ExplanationGenerator.py:84-93).
//
// The columns of R evolve independently (column c of Abar_l R only needs column c of R), so one CTA owns a block of 32
// columns of one sample's R for the WHOLE chain: that block lives in shared memory (double-buffered), the Abar_l planes
// stream through a second double buffer with cp.async (layer l+1 lands while layer l is multiplied), and R reaches HBM
// once, at the end.  The products run on the tensor cores as mma.sync.m16n8k8 TF32 with the fp32-faithful 3-pass split of
// the attention kernels (x = hi + lo, hi*hi in one accumulator, lo*hi + hi*lo in a second one, added once).  tcgen05 is the
// wrong tool at these sizes: S = 50 / 77 / 20 / 36 / 100 is below one 128-row UMMA tile (S >= 128 goes to the tcgen05
// GEMM in rules.cu).  Replaces one bmm launch per layer (12 + 12 per CLIP step) by one launch per tower, and the
// S x S x 4 bytes of R traffic per layer by none.
#include "mmx_common.cuh"

namespace mmx {
namespace chain {

constexpr int NB = 32;        // columns of R per CTA
constexpr int THREADS = 256;  // 8 warps

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

__host__ __device__ inline int lda_s(int S) { return round_up(S, 8) + 4; }   // [row][k] operand: stride = 4 (mod 8)
constexpr int LDR = NB + 8;                                                    // [k][n] operand: stride = 8 (mod 16)

inline size_t smem_bytes(int S) {
  const int Sp = round_up(S, 16);
  return sizeof(float) * ((size_t)2 * Sp * lda_s(S) + (size_t)2 * Sp * LDR);
}

// Abar: [L][B][S][ld] (layer stride `ls`, sample stride S*ld); R_out: [B][S][ld_out]
__global__ void __launch_bounds__(THREADS) rule6_chain_kernel(const float* __restrict__ Abar, long long ls, int ld,
                                                              float* __restrict__ R_out, int ld_out, int S, int L) {
  extern __shared__ float smem[];
  const int Sp = round_up(S, 16), LDA = lda_s(S);
  float* sA[2] = {smem, smem + (size_t)Sp * LDA};
  float* sR[2] = {smem + (size_t)2 * Sp * LDA, smem + (size_t)2 * Sp * LDA + (size_t)Sp * LDR};
  const int b = blockIdx.y, c0 = blockIdx.x * NB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const float* Ab = Abar + (long long)b * S * ld;

  // zero the Abar buffers once (pad rows / columns are never written afterwards), R = I on this column block
  for (int e = tid; e < 2 * Sp * LDA; e += THREADS) smem[e] = 0.f;
  for (int e = tid; e < Sp * LDR; e += THREADS) {
    const int i = e / LDR, c = e % LDR;
    sR[0][e] = (c < NB && i < S && i == c0 + c) ? 1.f : 0.f;
    sR[1][e] = 0.f;
  }
  __syncthreads();
  const int chunks = ld / 4;
  auto load_layer = [&](int l, float* dst) {
    const float* src = Ab + (long long)l * ls;
    for (int e = tid; e < S * chunks; e += THREADS) {
      const int r = e / chunks, c = (e % chunks) * 4;
      cp_async16(dst + r * LDA + c, src + (long long)r * ld + c);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_layer(0, sA[0]);
  const int mb = Sp / 16, tiles = mb * (NB / 8), ksteps = round_up(S, 8) / 8;
  int cur = 0;
  for (int l = 0; l < L; ++l) {
    if (l + 1 < L) {
      load_layer(l + 1, sA[(l + 1) & 1]);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();                                  // Abar_l landed; the previous layer's R writes are visible
    const float* A = sA[l & 1];
    const float* Rc = sR[cur];
    float* Rn = sR[cur ^ 1];
    for (int tile = warp; tile < tiles; tile += THREADS / 32) {
      const int m0 = (tile % mb) * 16, n0 = (tile / mb) * 8;
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, crs[4] = {0.f, 0.f, 0.f, 0.f};
      for (int ks = 0; ks < ksteps; ++ks) {
        const int k0 = ks * 8;
        uint32_t ahi[4], alo[4], bhi[2], blo[2];
        const float* pa = A + (m0 + g) * LDA + k0 + t;
        split_tf32(pa[0], ahi[0], alo[0]);
        split_tf32(pa[8 * LDA], ahi[1], alo[1]);
        split_tf32(pa[4], ahi[2], alo[2]);
        split_tf32(pa[8 * LDA + 4], ahi[3], alo[3]);
        const float* pb = Rc + (k0 + t) * LDR + n0 + g;
        split_tf32(pb[0], bhi[0], blo[0]);
        split_tf32(pb[4 * LDR], bhi[1], blo[1]);
        mma_tf32(crs, alo, bhi);
        mma_tf32(crs, ahi, blo);
        mma_tf32(acc, ahi, bhi);
      }
      // C fragment: (m0+g, n0+2t), (m0+g, n0+2t+1), (m0+g+8, n0+2t), (m0+g+8, n0+2t+1);  R_new = R_old + Abar R_old
      const int r0 = (m0 + g) * LDR + n0 + 2 * t, r1 = r0 + 8 * LDR;
      Rn[r0] = Rc[r0] + (acc[0] + crs[0]);
      Rn[r0 + 1] = Rc[r0 + 1] + (acc[1] + crs[1]);
      Rn[r1] = Rc[r1] + (acc[2] + crs[2]);
      Rn[r1 + 1] = Rc[r1 + 1] + (acc[3] + crs[3]);
    }
    cur ^= 1;
    __syncthreads();                                  // every warp is done with sA[l & 1] before layer l + 2 overwrites it
  }
  // this column block of R -> HBM (pad columns S .. ld_out-1 as zeros)
  const float* Rf = sR[cur];
  float* out = R_out + (long long)b * S * ld_out;
  for (int e = tid; e < S * NB; e += THREADS) {
    const int i = e / NB, c = e % NB;
    if (c0 + c < ld_out) out[(long long)i * ld_out + c0 + c] = (c0 + c < S) ? Rf[i * LDR + c] : 0.f;
  }
}

}  // namespace chain

bool rule6_chain_ok(int S, int ld, int ld_out, const float* Abar, const float* R_out) {
  return S >= 1 && S <= 128 && ld % 4 == 0 && ld >= S && ld_out >= S && aligned16(Abar) && chain::smem_bytes(S) <= 200 * 1024;
}

int rule6_chain(const float* Abar, long long layer_stride, int ld, float* R_out, int ld_out, int B, int S, int L, cudaStream_t st) {
  if (B == 0 || S == 0) return 0;
  MMX_REQUIRE(L >= 1, "at least one layer");
  MMX_REQUIRE(rule6_chain_ok(S, ld, ld_out, Abar, R_out), "rule-6 chain kernel: S <= 128, ld % 4 == 0, 16-byte aligned Abar");
  MMX_REQUIRE(layer_stride % 4 == 0 && ((long long)S * ld) % 4 == 0, "layer / sample strides must keep rows 16-byte aligned");
  const size_t smem = chain::smem_bytes(S);
  MMX_CHECK_CUDA(cudaFuncSetAttribute(chain::rule6_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int ncols = ld_out < round_up(S, chain::NB) ? ld_out : round_up(S, chain::NB);   // pad columns S .. ld_out-1 are zeroed too
  dim3 grid(cdiv(ncols, chain::NB), B);
  chain::rule6_chain_kernel<<<grid, chain::THREADS, smem, st>>>(Abar, layer_stride, ld, R_out, ld_out, S, L);
  MMX_LAUNCH_CHECK();
  return 0;
}

}  // namespace mmx

extern "C" int mmx_self_chain(const float* Abar, long long layer_stride, int ld, float* R_out, int ld_out, int B, int S, int L,
                              void* stream) {
  MMX_REQUIRE(Abar && R_out, "bad arguments");
  return mmx::rule6_chain(Abar, layer_stride, ld, R_out, ld_out, B, S, L, (cudaStream_t)stream);
}
