// bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands
// staged in shared memory by TMA with the 128-byte swizzle), persistent + warp specialised:
//   warp 0  : TMA producer            (one elected lane)
//   warp 1  : tcgen05.mma issuer      (one elected lane)
//   warp 2  : TMEM allocator
//   warps 4-7: epilogue (TMEM -> registers -> fused epilogue -> global)
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
//
// Replaces, for the VL-T5 hot path, every cuBLAS sgemm the reference reaches through nn.Linear
// (SURVEY.md §2.3 K1/K8/K9/K14): forward Y = X W^T (both operands K-major), dX = dY W (B MN-major),
// dW = dY^T X (both MN-major, split-K with fp32 red.global.add).
#pragma once
#include "common.cuh"

namespace vq {

enum GemmEpi : int {
  EPI_BF16 = 0,          // C(bf16) = alpha * acc
  EPI_RELU_BF16 = 1,     // C(bf16) = dropout(relu(acc))
  EPI_RESID_F32 = 2,     // C(f32)  = R(f32) + dropout(alpha * acc)
  EPI_ATOMIC_F32 = 3,    // C(f32) += alpha * acc (red.global.add; split-K capable)
  EPI_RELUBWD_BF16 = 4,  // C(bf16) = acc * (R(bf16) > 0 ? alpha : 0)   R = saved relu(+dropout) output
  EPI_F32 = 5,           // C(f32)  = alpha * acc
  EPI_BF16_ROWMASK = 6   // C(bf16) = alpha * acc for rows with (row % rowmod) < rowkeep, else skipped
};

struct GemmArgs {
  int epi;            // GemmEpi
  int M, N, K;        // C is MxN, contraction length K
  void* C;
  int ldc;            // elements
  const void* R;
  int ldr;            // elements
  float alpha;
  int splits;         // split-K factor (EPI_ATOMIC_F32 only)
  uint32_t drop_thr;  // p * 2^32 (0 = no dropout)
  float drop_inv_keep;
  uint32_t seed, site;
  int rowmod, rowkeep;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 256;

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmArgs p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_m = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int kblocks = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int kb_per_split = (kblocks + p.splits - 1) / p.splits;
  const int tiles_mn = tiles_m * tiles_n;
  const int total_work = tiles_mn * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int split = w / tiles_mn;
        const int t = w - split * tiles_mn;
        const int m_blk = t / tiles_n, n_blk = t - m_blk * tiles_n;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kblocks, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (!A_MN) {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
          } else {
#pragma unroll
            for (int j = 0; j < GEMM_BM / 64; ++j)
              tma_load_2d(sa + j * (GEMM_BK * 128), &tmA, &full_bar[stage], m_blk * GEMM_BM + j * 64, kb * GEMM_BK);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * GEMM_BK, n_blk * BN);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * (GEMM_BK * 128), &tmB, &full_bar[stage], n_blk * BN + j * 64, kb * GEMM_BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int astage = 0;
      uint32_t aphase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int split = w / tiles_mn;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kblocks, kb0 + kb_per_split);
        mbar_wait(&tempty_bar[astage], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + astage * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // K-major: 8-row groups 1024 B apart (SBO), +32 B per 16-element K step inside the 128 B swizzle row.
            // MN-major: 64-element MN atoms BK*128 B apart (LBO), 8-row K groups 1024 B apart (SBO),
            //           +2048 B per 16-row K step.
            const uint64_t adesc = A_MN ? umma_smem_desc_sw128(sa + k * 2048, GEMM_BK * 128, 1024)
                                        : umma_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? umma_smem_desc_sw128(sb + k * 2048, GEMM_BK * 128, 1024)
                                        : umma_smem_desc_sw128(sb + k * 32, 16, 1024);
            umma_f16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[astage]);   // accumulator complete -> epilogue
        if (++astage == 2) { astage = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ----------------------------------
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int astage = 0;
    uint32_t aphase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int split = w / tiles_mn;
      const int t = w - split * tiles_mn;
      const int m_blk = t / tiles_n, n_blk = t - m_blk * tiles_n;
      const int kb0 = split * kb_per_split;
      const bool has_k = kb0 < kblocks;
      mbar_wait(&tfull_bar[astage], aphase);
      tc_fence_after();
      const int row = m_blk * GEMM_BM + q * 32 + lane;
      const bool row_ok = row < p.M && has_k &&
                          (p.rowmod <= 0 || (row % p.rowmod) < p.rowkeep);
      const uint32_t t_base = tmem_base + astage * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_base + c * 32, r);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (row_ok && col0 < p.N) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
          const int ncols = min(32, p.N - col0);  // multiple of 8 (host-checked)
          const size_t coff = (size_t)row * p.ldc + col0;
          const int epi = p.epi;
          if (epi == EPI_BF16 || epi == EPI_BF16_ROWMASK) {
            __nv_bfloat16* C = reinterpret_cast<__nv_bfloat16*>(p.C) + coff;
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (g * 8 < ncols) {
                uint4 o;
                o.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
                o.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
                o.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
                o.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
                *reinterpret_cast<uint4*>(C + g * 8) = o;
              }
          } else if (epi == EPI_RELU_BF16) {
            __nv_bfloat16* C = reinterpret_cast<__nv_bfloat16*>(p.C) + coff;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float x = fmaxf(v[i], 0.0f);
              if (p.drop_thr) x *= vq_dropout_scale(p.seed, p.site, (uint64_t)row * p.N + col0 + i, p.drop_thr, p.drop_inv_keep);
              v[i] = x;
            }
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (g * 8 < ncols) {
                uint4 o;
                o.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
                o.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
                o.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
                o.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
                *reinterpret_cast<uint4*>(C + g * 8) = o;
              }
          } else if (epi == EPI_RESID_F32) {
            float* C = reinterpret_cast<float*>(p.C) + coff;
            const float* R = reinterpret_cast<const float*>(p.R) + (size_t)row * p.ldr + col0;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              if (g * 4 < ncols) {
                float4 rr = *reinterpret_cast<const float4*>(R + g * 4);
                float a0 = v[g * 4 + 0], a1 = v[g * 4 + 1], a2 = v[g * 4 + 2], a3 = v[g * 4 + 3];
                if (p.drop_thr) {
                  const uint64_t e = (uint64_t)row * p.N + col0 + g * 4;
                  a0 *= vq_dropout_scale(p.seed, p.site, e + 0, p.drop_thr, p.drop_inv_keep);
                  a1 *= vq_dropout_scale(p.seed, p.site, e + 1, p.drop_thr, p.drop_inv_keep);
                  a2 *= vq_dropout_scale(p.seed, p.site, e + 2, p.drop_thr, p.drop_inv_keep);
                  a3 *= vq_dropout_scale(p.seed, p.site, e + 3, p.drop_thr, p.drop_inv_keep);
                }
                rr.x += a0; rr.y += a1; rr.z += a2; rr.w += a3;
                *reinterpret_cast<float4*>(C + g * 4) = rr;
              }
          } else if (epi == EPI_ATOMIC_F32) {
            float* C = reinterpret_cast<float*>(p.C) + coff;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              if (g * 4 < ncols) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(C + g * 4), "f"(v[g * 4 + 0]),
                             "f"(v[g * 4 + 1]), "f"(v[g * 4 + 2]), "f"(v[g * 4 + 3])
                             : "memory");
              }
          } else if (epi == EPI_RELUBWD_BF16) {
            __nv_bfloat16* C = reinterpret_cast<__nv_bfloat16*>(p.C) + coff;
            const __nv_bfloat16* H = reinterpret_cast<const __nv_bfloat16*>(p.R) + (size_t)row * p.ldr + col0;
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (g * 8 < ncols) {
                uint4 h = *reinterpret_cast<const uint4*>(H + g * 8);
                const uint32_t hh[4] = {h.x, h.y, h.z, h.w};
                uint32_t oo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float2 hv = unpack_bf16(hh[j]);
                  // v already carries alpha (= 1/keep of the inner dropout)
                  oo[j] = pack_bf16(hv.x > 0.0f ? v[g * 8 + 2 * j] : 0.0f, hv.y > 0.0f ? v[g * 8 + 2 * j + 1] : 0.0f);
                }
                *reinterpret_cast<uint4*>(C + g * 8) = make_uint4(oo[0], oo[1], oo[2], oo[3]);
              }
          } else {  // EPI_F32
            float* C = reinterpret_cast<float*>(p.C) + coff;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              if (g * 4 < ncols)
                *reinterpret_cast<float4*>(C + g * 4) = make_float4(v[g * 4 + 0], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[astage]);
      if (++astage == 2) { astage = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace vq
