// GEMM backends shared declarations.
#pragma once
#include "mmx_common.cuh"

namespace mmx {

struct GemmEpilogue {
  const float* bias = nullptr;      // [N]: added first
  const float* pre = nullptr;       // [M,N]: multiply by act'(pre)  (dgrad through an activation)
  int ldpre = 0;
  const float* residual = nullptr;  // [M,N]: added last
  int ldres = 0;
  float* C_act = nullptr;           // [M,N] (row stride ldc): also store act(C)
  int act = MMX_ACT_NONE;
  const int* m_dev = nullptr;       // optional device-resident row count (<= M): M becomes an upper bound (ragged batches)
};

// C[M,N] = A[M,K] * Bt[N,K]^T, both operands K-major.
int gemm_nt_simt(const float* A, int lda, const float* Bt, int ldb, float* C, int ldc, int M, int N, int K,
                 const GemmEpilogue& ep, cudaStream_t st);

// Backend dispatch: tcgen05 3xTF32 when the backend is enabled and the shape qualifies (the rule depends on N, K and
// alignment only - never on M - so one sample alone and inside a batch take the same arithmetic path), else fp32 FFMA.
int gemm_nt(const float* A, int lda, const float* Bt, int ldb, float* C, int ldc, int M, int N, int K,
            const GemmEpilogue& ep, cudaStream_t st);

int gemm_tc_available();
int gemm_tc_tile_n(int bn);   // forced tcgen05 tile width: bn < 0 reads it, 0 = automatic, 128/144/160 force

// Tensor-core product for the relevancy updates (no bias / activation): C = residual + A * Bt^T with N possibly not a
// multiple of 4 as long as every row stride covers round_up(N, 4) columns (the pad columns receive zeros).  Returns
// taken = false when the shapes / alignment / backend do not qualify.
int gemm_nt_tc_rule(const float* A, int lda, const float* Bt, int ldb, const float* residual, int ldres, float* C, int ldc,
                    int M, int N, int K, cudaStream_t st, bool* taken);
int gemm_backend();

}  // namespace mmx
