// Host-side interface of the tcgen05 GEMM (used by the engine and by the C-ABI).
#pragma once
#include "gemm_tcgen05.cuh"

namespace vq {

struct GemmOperand {
  const void* ptr;  // bf16
  int ld;           // pitch in elements of the stored 2-D tensor
  bool mn_major;    // false: stored [MN rows, K contiguous]; true: stored [K rows, MN contiguous]
};

// C[M,N] (+)= A[M,K] * B[N,K]^T with the epilogue in args.epi. force_bn = 0 picks the N tile by occupancy.
int gemm_bf16(const GemmOperand& A, const GemmOperand& B, GemmArgs args, int force_bn, cudaStream_t stream);
void gemm_tmap_cache_clear();
int num_sms();

}  // namespace vq
