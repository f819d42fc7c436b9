// bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands
// staged in shared memory by TMA with the 128-byte swizzle), persistent + warp specialised:
//   warp 0  : TMA producer            (one elected lane)
//   warp 1  : tcgen05.mma issuer      (one elected lane)
//   warp 2  : TMEM allocator
//   warps 4-11: epilogue (TMEM -> registers -> fused math -> swizzled smem transposition -> coalesced global), two warps
//              per TMEM lane quarter working on alternate 32-column chunks
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
  EPI_RELU_BF16 = 1,     // C(bf16) = dropout(relu(alpha * acc)); if R != null also writes the sign bitmask (one u32 per row and
                         // 32 columns, pitch ldr words) that EPI_RELUBWD_BF16 consumes
  EPI_RESID_F32 = 2,     // C(f32)  = R(f32) + dropout(alpha * acc)
  EPI_ATOMIC_F32 = 3,    // C(f32) += alpha * acc (red.global.add; split-K capable)
  EPI_RELUBWD_BF16 = 4,  // C(bf16) = mask bit ? alpha * acc : 0      R = u32 bitmask written by EPI_RELU_BF16 (16x less traffic than
                         // re-reading the saved activations)
  EPI_F32 = 5,           // C(f32)  = alpha * acc
  EPI_ARGMAX = 6         // no C matrix: per row the first maximum of alpha * acc over this tile's columns, straight from the fp32
                         // accumulators. C = float [M, ldc] values, R = int32 [M, ldr] column indices, slot = n_block * 2 + (index of
                         // the epilogue warp among the two sharing a TMEM lane quarter); needs 256-wide tiles (force_bn 256 / 512).
                         // Greedy decoding never materialises the [B, 32200] logits and the argmax is not taken on bf16-rounded
                         // values (two fp32 logits 0.01 apart round to the same bf16 above 2.0).
};

// row tail 2: the rest of VisualEmbedding.forward (modeling_t5_our.py:93-143) on the rows of feats @ Wf^T just written to C
struct VisTail {
  const float *boxes, *bf, *wf, *Wp, *bp, *wp, *img_emb, *shared;
  float* x;                      // [B, S, 768]: rows [L, L + N) of every batch element are written
  int V, N, S, L;
  uint32_t drop_thr; float drop_inv_keep; uint32_t drop_seed;
};

struct GemmArgs {
  int epi;            // GemmEpi
  int M, N, K;        // C is MxN, contraction length K
  void* C;
  int ldc;            // elements
  const void* R;
  int ldr;            // elements
  float alpha;
  int splits;         // split-K factor: EPI_ATOMIC_F32 (the slices meet in C through red.global.add), or EPI_F32 with
                      // split_stride > 0: slice s stores its partial product to C + s * split_stride (elements) and the
                      // consumer adds the slabs in a fixed order — the deterministic form, used where the sum is a forward
                      // activation (decoder FFN-out: the next RMSNorm kernel forms residual + dropout(sum of slabs))
  long long split_stride;
  uint32_t drop_thr;  // 16-bit keep threshold (0 = no dropout), see vq_dropout_pair
  float drop_inv_keep;
  uint32_t seed, site;  // seed = per-launch dropout key (already mixed with the site id); site is informational
  // CTA-pair kernel only — row tail: every pair OWNS whole 256-row blocks (it computes all N tiles of a block back to back), so
  // once the block's last tile has been written its epilogue warps can run a row-wise follow-up on the finished rows while they
  // are still in L2, instead of a separate kernel launch:  tail = 1: T5 RMSNorm of the fp32 rows just written (N must be 768):
  // tail_out(bf16)[r, :] = C[r, :] * rsqrt(mean(C[r, :]^2) + tail_eps) * tail_w   — the norm that feeds the NEXT sub-layer's GEMM.
  // tail = 2: VisualEmbedding tail (C = feats @ Wf^T, 768 wide): x[b, L + n] = RMSNorm(C + bf) wf + RMSNorm([box, area] Wp^T + bp) wp
  // + img_order_embedding[0] + shared[V - 1 - n], with the embedding dropout (VisTail vt) — the whole VisualEmbedding in one launch.
  int tail;
  VisTail vt;
  const float* tail_w;
  void* tail_out;
  int tail_ld;          // elements
  float tail_eps;
  int sm_limit;         // host side only: launch on at most this many SMs (0 = all). The decoder's weight-gradient GEMMs run on a
                        // second stream beside a chain of small latency-bound kernels; a persistent 148-CTA launch with 225 KB of
                        // shared memory per CTA holds EVERY SM until it ends, and the chain's next GEMM waits for it (measured:
                        // 10 us per cross-attention backward, profiles/r02_step_trace_decoder.txt)
  uint32_t* sched;      // CTA-pair kernel only: {next work item, pairs finished} counters of a DYNAMIC tile schedule (null = static
                        // round robin). A pair that becomes resident late — its SMs were held by a collective's CTAs or by another
                        // stream's kernel — then finds the work list drained instead of running its whole static share afterwards.
};

// A group of up to GEMM_GROUP_MAX independent problems that share K, the operand majors and the epilogue, computed by ONE launch
// of the CTA-pair kernel (its work list is the concatenation of the problems' 256 x 256 tiles; no split-K, no row tail).
// Made for the decoder's weight gradients: six dW = dY^T X per layer with K = B*T = 1 600 rows, 18-72 tiles each — alone each
// is a launch whose pipeline fill and drain outweigh its 25 k-blocks (25 % of the tensor peak, measured), together they are
// 126 pair tiles = 1.7 waves of one persistent launch.
constexpr int GEMM_GROUP_MAX = 8;
struct GemmGroup {
  int n;
  int M[GEMM_GROUP_MAX], N[GEMM_GROUP_MAX], ldc[GEMM_GROUP_MAX], tiles_n[GEMM_GROUP_MAX];
  int tile_start[GEMM_GROUP_MAX + 1];     // prefix sums of the problems' pair-tile counts
  void* C[GEMM_GROUP_MAX];
};
struct GemmGroupMaps { CUtensorMap a[GEMM_GROUP_MAX], b[GEMM_GROUP_MAX]; };

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
#ifndef GEMM_EPI_WARPS_N
#define GEMM_EPI_WARPS_N 8
#endif
constexpr int GEMM_EPI_WARPS = GEMM_EPI_WARPS_N;            // 8 or 12: 2 or 3 warps per TMEM lane quarter, taking every 2nd / 3rd 32-column chunk.
                                                             // Measured (profiles/r01_summary.md): 12 warps (512 threads, one smem stage less) make the
                                                             // fp32-residual epilogue 6 % faster and every MMA-bound shape 3-5 % slower; 8 it is.
constexpr int GEMM_EPI_GROUPS = GEMM_EPI_WARPS / 4;
constexpr int GEMM_THREADS = 128 + GEMM_EPI_WARPS * 32;
#ifndef GEMM_MAXREG
#define GEMM_MAXREG 128   // 384 threads x 128 = 48 K registers: leaves room on the SM for a memory-bound kernel of another stream (168 measured no faster)
#endif
constexpr int EPI_TILE_BYTES = 4096;                         // one 32x32 fp32 (or 32x32 bf16 in half of it) tile per epilogue warp
constexpr int GEMM_SMEM_MAX = 232448;                        // opt-in maximum per CTA on sm_100
constexpr int gemm_stages(int stage_bytes, int max_stages) {  // as many pipeline stages as fit beside the epilogue staging tiles
  const int fit = (GEMM_SMEM_MAX - 1024 - 256 - GEMM_EPI_WARPS * EPI_TILE_BYTES) / stage_bytes;
  return fit < max_stages ? fit : max_stages;
}

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = gemm_stages(STAGE_BYTES, (BN == 256) ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + GEMM_EPI_WARPS * EPI_TILE_BYTES;
};

VQ_DEVINL void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
VQ_DEVINL uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// One 128 x BN accumulator tile: this warp owns TMEM lanes [32q, 32q+32) (= tile rows) and the 32-column chunks
// c = half, half + GEMM_EPI_GROUPS, ... (`half` = index of the warp among those sharing its lane quarter) tcgen05.ld 32x32b gives thread = row, registers = 32 consecutive columns; a global access from that
// layout is 32 row-strided 16-byte requests per instruction. The fused math therefore runs in the row layout, the result is
// transposed through a private XOR-swizzled smem tile (conflict-free 16-byte writes and reads, no padding), and global
// memory is touched with 4 (bf16: 8) full rows of 128 (64) contiguous bytes per instruction. Extra operands (residual R,
// saved activations H) are fetched with the same coalesced pattern before the accumulator wait.
template <int EPI, int BN>
VQ_DEVINL void gemm_epilogue_tile(const GemmArgs& p, uint32_t t_base, uint32_t stg, int row_base, int n_base, int half, int lane,
                                  bool has_k, uint64_t* tfull, uint32_t aphase) {
  if constexpr (EPI == EPI_ARGMAX) {
    const int rlim = has_k ? p.M - row_base : 0;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    mbar_wait(tfull, aphase);
    tc_fence_after();
#pragma unroll 1
    for (int c = half; c < BN / 32; c += GEMM_EPI_GROUPS) {
      const int col0 = n_base + c * 32;
      if (col0 >= p.N) break;
      uint32_t r[32];
      tmem_ld_32x32(t_base + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float v = __uint_as_float(r[i]) * p.alpha;
        if (col0 + i < p.N && v > best) { best = v; bi = col0 + i; }   // ascending columns + strict '>' = first maximum
      }
    }
    if (lane < rlim) {
      const int slot = (n_base / BN) * GEMM_EPI_GROUPS + half;
      reinterpret_cast<float*>(p.C)[(size_t)(row_base + lane) * p.ldc + slot] = best;
      reinterpret_cast<int*>(const_cast<void*>(p.R))[(size_t)(row_base + lane) * p.ldr + slot] = bi;
    }
    return;
  }
  constexpr bool OUT_BF16 = (EPI == EPI_BF16 || EPI == EPI_RELU_BF16 || EPI == EPI_RELUBWD_BF16);
  constexpr int ESZ = OUT_BF16 ? 2 : 4;          // bytes per output element
  constexpr int RPI = OUT_BF16 ? 8 : 4;          // rows one store instruction covers (4 or 8 lanes x 16 B per row)
  constexpr int NIT = 32 / RPI;                  // store instructions per 32-row chunk
  constexpr int ROWB = 32 * ESZ;                 // bytes of one staged row (32 columns)
  const int N = p.N;
  const float alpha = p.alpha;
  // Everything that does not change from chunk to chunk is computed once per tile: the number of rows of this warp's 32-row
  // slab that exist, this lane's byte pointers into C / R for its first row (later rows and chunks are reached by adding
  // strides) and its shared-memory staging addresses. The staging tile is at least 128-byte aligned, so the XOR swizzle
  // of a 16-byte piece index can be applied to the byte address.
  const int rlim = has_k ? p.M - row_base : 0;                      // row r of the slab is written iff r < rlim
  const int rr0 = OUT_BF16 ? (lane >> 2) : (lane >> 3);             // row inside a store instruction
  const int piece = OUT_BF16 ? (lane & 3) : (lane & 7);             // 16-byte piece of that row
  const int pcol = n_base + piece * (16 / ESZ);                     // + c * 32 = first column of this lane's piece
  char* cbase = reinterpret_cast<char*>(p.C) + ((size_t)(row_base + rr0) * p.ldc + pcol) * ESZ;
  const size_t cstep = (size_t)RPI * p.ldc * ESZ;
  const uint32_t sts_base = stg + lane * ROWB + ((OUT_BF16 ? ((lane >> 1) & 3) : (lane & 7)) << 4);   // ^ (j << 4) per piece j
  // fp32: rows it * 4 + rr0 have swizzle key (row & 7) = rr0 + 4 * (it & 1), i.e. odd `it` flip bit 2 of the piece index
  const uint32_t lds_base = stg + rr0 * ROWB + ((OUT_BF16 ? (piece ^ ((rr0 >> 1) & 3)) : (piece ^ rr0)) << 4);
  // extra-operand registers are double-buffered in time: the operand of this warp's next chunk is requested right
  // after the accumulator of chunk c has been read, so its global-memory latency overlaps the math and stores of chunk c
  float4 rres[8];
  uint32_t rmask = 0;   // ReLU-backward: bit i = activation (row = this lane's row, column col0 + i) was positive and kept
  const char* rbase = nullptr;
  size_t rstep = 0;
  if (EPI == EPI_RESID_F32) {
    rbase = reinterpret_cast<const char*>(p.R) + ((size_t)(row_base + rr0) * p.ldr + pcol) * 4;
    rstep = (size_t)4 * p.ldr * 4;
  }
  if (EPI == EPI_RELU_BF16 || EPI == EPI_RELUBWD_BF16)   // one mask word per (row = this lane's accumulator row, chunk)
    rbase = reinterpret_cast<const char*>(p.R) + ((size_t)(row_base + lane) * p.ldr + (n_base >> 5)) * 4;
  auto prefetch = [&](int cc) {
    const bool ok = cc < BN / 32 && n_base + cc * 32 < N;
    if (EPI == EPI_RESID_F32) {
      const bool colok = ok && pcol + cc * 32 < N;
      const char* rp = rbase + cc * 128;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        rres[it] = (colok && it * 4 + rr0 < rlim) ? *reinterpret_cast<const float4*>(rp) : make_float4(0.f, 0.f, 0.f, 0.f);
        rp += rstep;
      }
    }
    if (EPI == EPI_RELUBWD_BF16) rmask = (ok && lane < rlim) ? *reinterpret_cast<const uint32_t*>(rbase + cc * 4) : 0u;
  };
  prefetch(half);
  bool waited = false;
#pragma unroll 1
  for (int c = half; c < BN / 32; c += GEMM_EPI_GROUPS) {
    const int col0 = n_base + c * 32;
    if (col0 >= N) break;
    float4 cres[8];
    if (EPI == EPI_RESID_F32) {
#pragma unroll
      for (int it = 0; it < 8; ++it) cres[it] = rres[it];
    }
    const uint32_t cmask = rmask;
    if (!waited) {
      mbar_wait(tfull, aphase);
      tc_fence_after();
      waited = true;
    }
    uint32_t r[32];
    tmem_ld_32x32(t_base + c * 32, r);
    tmem_ld_wait();
    prefetch(c + GEMM_EPI_GROUPS);
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * alpha;
    if (EPI == EPI_RELU_BF16) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if ((EPI == EPI_RELU_BF16 || EPI == EPI_RESID_F32) && p.drop_thr) {
      const uint32_t pi = (uint32_t)(((size_t)(row_base + lane) * N + col0) >> 1);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float s0, s1;
        vq_dropout_pair(p.seed, pi + j, p.drop_thr, p.drop_inv_keep, s0, s1);
        v[2 * j] *= s0;
        v[2 * j + 1] *= s1;
      }
    }
    if (EPI == EPI_RELU_BF16 && p.R) {
      // sign bitmask of the stored activations for the backward pass (dropped elements count as zero)
      uint32_t m = 0;
#pragma unroll
      for (int i = 0; i < 32; ++i) m |= (v[i] > 0.f ? 1u : 0u) << i;
      if (lane < rlim) *reinterpret_cast<uint32_t*>(const_cast<char*>(rbase) + c * 4) = m;
    }
    if (EPI == EPI_RELUBWD_BF16) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = ((cmask >> i) & 1u) ? v[i] : 0.f;
    }
    // stage the 32 x 32 chunk (thread = row) with its 16-byte pieces XOR-swizzled, read it back as full rows
    if (OUT_BF16) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        sts128(sts_base ^ (j << 4), pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
               pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts128(sts_base ^ (j << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
               __float_as_uint(v[4 * j + 3]));
    }
    __syncwarp();
    const bool colok = pcol + c * 32 < N;
    char* cp = cbase + c * (32 * ESZ);
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const uint4 u = lds128((OUT_BF16 || !(it & 1) ? lds_base : (lds_base ^ 64u)) + it * (RPI * ROWB));
      if (colok && it * RPI + rr0 < rlim) {
        if (EPI == EPI_RESID_F32) {
          float4 o = make_float4(__uint_as_float(u.x) + cres[it].x, __uint_as_float(u.y) + cres[it].y, __uint_as_float(u.z) + cres[it].z,
                                 __uint_as_float(u.w) + cres[it].w);
          *reinterpret_cast<float4*>(cp) = o;
        } else if (EPI == EPI_ATOMIC_F32) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp), "f"(__uint_as_float(u.x)), "f"(__uint_as_float(u.y)),
                       "f"(__uint_as_float(u.z)), "f"(__uint_as_float(u.w))
                       : "memory");
        } else {
          *reinterpret_cast<uint4*>(cp) = u;
        }
      }
      cp += cstep;
    }
    __syncwarp();
  }
  if (!waited) {
    mbar_wait(tfull, aphase);
    tc_fence_after();
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __maxnreg__(GEMM_MAXREG)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmArgs p) {
  vq_pdl_trigger();
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // the 128-byte swizzle needs 1024-byte aligned stage bases; do not rely on the attribute alone
  const uint32_t raw_u32 = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_u32 + 1023u) & ~1023u) - raw_u32);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* epi_stage = smem + STAGES * Cfg::STAGE_BYTES + 256;

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
      mbar_init(&tempty_bar[i], GEMM_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  vq_pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched only below

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
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;  // which alternate 32-column chunks it handles
    const uint32_t stg = smem_u32(epi_stage) + (warp - 4) * EPI_TILE_BYTES;
    int astage = 0;
    uint32_t aphase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int split = w / tiles_mn;
      const int t = w - split * tiles_mn;
      const int m_blk = t / tiles_n, n_blk = t - m_blk * tiles_n;
      const bool has_k = split * kb_per_split < kblocks;
      const int row_base = m_blk * GEMM_BM + q * 32;
      const uint32_t t_base = tmem_base + astage * BN + ((uint32_t)(q * 32) << 16);
      uint64_t* tf = &tfull_bar[astage];
      if (p.epi == EPI_F32 && p.split_stride) {
        GemmArgs pg = p;
        pg.C = reinterpret_cast<float*>(p.C) + (size_t)split * p.split_stride;
        gemm_epilogue_tile<EPI_F32, BN>(pg, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase);
      } else
      switch (p.epi) {
        case EPI_BF16: gemm_epilogue_tile<EPI_BF16, BN>(p, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase); break;
        case EPI_RELU_BF16: gemm_epilogue_tile<EPI_RELU_BF16, BN>(p, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase); break;
        case EPI_RESID_F32: gemm_epilogue_tile<EPI_RESID_F32, BN>(p, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase); break;
        case EPI_ATOMIC_F32: gemm_epilogue_tile<EPI_ATOMIC_F32, BN>(p, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase); break;
        case EPI_RELUBWD_BF16: gemm_epilogue_tile<EPI_RELUBWD_BF16, BN>(p, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase); break;
        case EPI_ARGMAX: gemm_epilogue_tile<EPI_ARGMAX, BN>(p, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase); break;
        default: gemm_epilogue_tile<EPI_F32, BN>(p, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase); break;
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

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs computes 256 x 256 tiles. Each CTA stages its own 128 rows
// of A and its half (128 of the 256 columns) of B per k-block, the leader CTA's elected lane issues one M = 256 MMA that
// reads both CTAs' shared memory and writes rows 0-127 to the leader's TMEM and rows 128-255 to the peer's. Compared with
// the single-CTA 128 x 256 tile this cuts shared-memory traffic per FLOP by a third (TMA writes + MMA reads: 128 instead of
// 192 B/cycle/SM at full tensor rate — the single-CTA kernel saturates at 2/3 tensor-pipe utilisation, profiles/r01_*).
//   full[s]   (leader's)        : 1 arrival (leader producer's expect_tx for BOTH CTAs' bytes) + the bytes of both CTAs' loads
//   empty[s]  (one per CTA)     : tcgen05.commit multicast to both CTAs when the MMAs that read stage s are done
//   tfull[a]  (one per CTA)     : tcgen05.commit multicast when a 256 x 256 accumulator is complete
//   tempty[a] (leader's)        : 16 arrivals = 8 epilogue warps of each CTA (the peer arrives remotely)
// ---------------------------------------------------------------------------------------------------------------------
// Row tail of the CTA-pair kernel (GemmArgs::tail == 1): T5 RMSNorm (hf5.5 modeling_t5.py:46-68) of the CTA's 128 freshly written
// fp32 rows [row0, row0 + 128) of C (768 wide), one warp per row, lane l holds the float4 chunks {l, l + 32, ..., l + 160}.
VQ_DEVINL void gemm_tail_rmsnorm(const GemmArgs& p, int row0, int ew, int lane) {
  constexpr int CH = 768 / 4 / 32;
  float4 wv[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) wv[j] = *reinterpret_cast<const float4*>(p.tail_w + (lane + 32 * j) * 4);
  const float* C = reinterpret_cast<const float*>(p.C);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.tail_out);
  for (int r = row0 + ew; r < row0 + GEMM_BM && r < p.M; r += GEMM_EPI_WARPS) {
    float4 v[CH];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      v[j] = __ldcg(reinterpret_cast<const float4*>(C + (size_t)r * p.ldc + (lane + 32 * j) * 4));   // L2: written by other warps of this CTA
      ss += v[j].x * v[j].x; ss += v[j].y * v[j].y; ss += v[j].z * v[j].z; ss += v[j].w * v[j].w;   // same order as rmsnorm_fwd_kernel
    }
    const float rstd = rsqrtf(warp_sum(ss) / 768.f + p.tail_eps);
#pragma unroll
    for (int j = 0; j < CH; ++j)
      *reinterpret_cast<uint2*>(out + (size_t)r * p.tail_ld + (lane + 32 * j) * 4) =
          make_uint2(pack_bf16(v[j].x * rstd * wv[j].x, v[j].y * rstd * wv[j].y), pack_bf16(v[j].z * rstd * wv[j].z, v[j].w * rstd * wv[j].w));
  }
}

// Row tail 2 (see GemmArgs): one warp per row, the same arithmetic in the same order as vis_embed_fwd_kernel (elementwise.cu),
// but in three passes over the row's six float4 chunks (two for the statistics, one that recomputes and writes) so that only a
// handful of values are live at a time: this code shares the GEMM kernel's 128-register budget.
VQ_DEVINL void gemm_tail_visual(const GemmArgs& p, int row0, int ew, int lane) {
  constexpr int CH = 768 / 4 / 32;
  const VisTail& a = p.vt;
  const float* C = reinterpret_cast<const float*>(p.C);
  const int rend = min(row0 + GEMM_BM, p.M);
  for (int r = row0 + ew; r < rend; r += GEMM_EPI_WARPS) {
    const int b = r / a.N, n = r - b * a.N;
    const float* crow = C + (size_t)r * p.ldc;
    const float4 bx = *reinterpret_cast<const float4*>(a.boxes + (size_t)r * 4);
    const float p5[5] = {bx.x, bx.y, bx.z, bx.w, (bx.w - bx.z) * (bx.y - bx.x)};   // area as the reference writes it (:78-90)
    auto feat4 = [&](int j, float (&u)[4]) {        // feats Wf^T + bf
      const int c0 = (lane + 32 * j) * 4;
      const float4 c = __ldcg(reinterpret_cast<const float4*>(crow + c0));
      const float4 t = *reinterpret_cast<const float4*>(a.bf + c0);
      u[0] = c.x + t.x; u[1] = c.y + t.y; u[2] = c.z + t.z; u[3] = c.w + t.w;
    };
    auto pos4 = [&](int j, float (&u)[4]) {         // [box, area] Wp^T + bp
      const int c0 = (lane + 32 * j) * 4;
      const float4 t = *reinterpret_cast<const float4*>(a.bp + c0);
      const float tb[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float* wrow = a.Wp + (size_t)(c0 + i) * 5;
        float acc = tb[i];
#pragma unroll
        for (int k = 0; k < 5; ++k) acc += wrow[k] * p5[k];
        u[i] = acc;
      }
    };
    float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
    for (int j = 0; j < CH; ++j) {
      float u[4];
      feat4(j, u);
#pragma unroll
      for (int i = 0; i < 4; ++i) s1 += u[i] * u[i];
    }
#pragma unroll 1
    for (int j = 0; j < CH; ++j) {
      float u[4];
      pos4(j, u);
#pragma unroll
      for (int i = 0; i < 4; ++i) s2 += u[i] * u[i];
    }
    const float rstd1 = rsqrtf(warp_sum(s1) / 768.f + p.tail_eps), rstd2 = rsqrtf(warp_sum(s2) / 768.f + p.tail_eps);
    const size_t orow = (size_t)b * a.S + a.L + n;
#pragma unroll 1
    for (int j = 0; j < CH; ++j) {
      const int c0 = (lane + 32 * j) * 4;
      float u[4], v[4], out[4];
      feat4(j, u);
      pos4(j, v);
      const float4 wf = *reinterpret_cast<const float4*>(a.wf + c0), wp = *reinterpret_cast<const float4*>(a.wp + c0);
      const float4 im = *reinterpret_cast<const float4*>(a.img_emb + c0);
      const float4 sh = *reinterpret_cast<const float4*>(a.shared + (size_t)(a.V - 1 - n) * 768 + c0);
      const float wfv[4] = {wf.x, wf.y, wf.z, wf.w}, wpv[4] = {wp.x, wp.y, wp.z, wp.w};
      const float imv[4] = {im.x, im.y, im.z, im.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        out[i] = u[i] * rstd1 * wfv[i];
        out[i] += v[i] * rstd2 * wpv[i];
        out[i] += imv[i] + shv[i];
      }
      if (a.drop_thr) {
        const uint32_t pi = (uint32_t)((orow * 768) >> 1) + (uint32_t)(lane + 32 * j) * 2u;
        float d0, d1, d2, d3;
        vq_dropout_pair(a.drop_seed, pi, a.drop_thr, a.drop_inv_keep, d0, d1);
        vq_dropout_pair(a.drop_seed, pi + 1, a.drop_thr, a.drop_inv_keep, d2, d3);
        out[0] *= d0; out[1] *= d1; out[2] *= d2; out[3] *= d3;
      }
      *reinterpret_cast<float4*>(a.x + orow * 768 + c0) = make_float4(out[0], out[1], out[2], out[3]);
    }
  }
}

constexpr int GEMM2_BN = 256;
struct Gemm2Cfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;              // 128 rows of A
  static constexpr int B_BYTES = (GEMM2_BN / 2) * GEMM_BK * 2;       // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;              // 32 KB
  static constexpr int STAGES = gemm_stages(STAGE_BYTES, 6);
  static constexpr int TMEM_COLS = 2 * GEMM2_BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + GEMM_EPI_WARPS * EPI_TILE_BYTES;
};

template <bool A_MN, bool B_MN, bool GROUPED>
VQ_DEVINL void gemm_pair_body(const CUtensorMap* tmA_, const CUtensorMap* tmB_, const GemmArgs& p, const GemmGroup* grp) {
  vq_pdl_trigger();
  using Cfg = Gemm2Cfg;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BN = GEMM2_BN;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_u32 + 1023u) & ~1023u) - raw_u32);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  // dynamic schedule: ring of work-item indices, filled by the leader's producer lane in BOTH CTAs
  constexpr int SCHED_DEPTH = 4;
  uint64_t* sfull_bar = reinterpret_cast<uint64_t*>(tmem_holder + 2);   // [SCHED_DEPTH] per CTA: slot written
  uint64_t* sempty_bar = sfull_bar + SCHED_DEPTH;                       // [SCHED_DEPTH] leader's: slot read by all 18 consumers
  uint32_t* sched_w = reinterpret_cast<uint32_t*>(sempty_bar + SCHED_DEPTH);
  uint8_t* epi_stage = smem + STAGES * Cfg::STAGE_BYTES + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();        // 0 = leader
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  const int tiles_m = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const int tiles_n = (p.N + BN - 1) / BN;
  const int kblocks = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int kb_per_split = GROUPED ? kblocks : (kblocks + p.splits - 1) / p.splits;
  const int tiles_mn = GROUPED ? grp->tile_start[grp->n] : tiles_m * tiles_n;
  // work items: one (split, m, n) tile each — or, with a row tail, one 256-row block each = `run` consecutive tiles (n fastest);
  // grouped: one tile of one problem each
  const int run = (!GROUPED && p.tail) ? tiles_n : 1;
  const int total_work = GROUPED ? tiles_mn : (p.tail ? tiles_m : tiles_mn * p.splits);
  // work index -> (problem, split, tile coordinates)
  auto decode = [&](int w, int& g, int& split, int& m_blk, int& n_blk) {
    if (GROUPED) {
      g = 0;
      while (w >= grp->tile_start[g + 1]) ++g;
      const int t = w - grp->tile_start[g];
      m_blk = t / grp->tiles_n[g]; n_blk = t - m_blk * grp->tiles_n[g];
      split = 0;
    } else {
      g = 0;
      split = w / tiles_mn;
      const int t = w - split * tiles_mn;
      m_blk = t / tiles_n; n_blk = t - m_blk * tiles_n;
    }
  };

  if (warp == 0 && lane == 0) {
    const int nmaps = GROUPED ? grp->n : 1;
    for (int g = 0; g < nmaps; ++g) {
      tma_prefetch_desc(tmA_ + g);
      tma_prefetch_desc(tmB_ + g);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * GEMM_EPI_WARPS);
    }
    for (int i = 0; i < SCHED_DEPTH; ++i) {
      mbar_init(&sfull_bar[i], 1);
      mbar_init(&sempty_bar[i], 2 + 2 * GEMM_EPI_WARPS);   // leader MMA lane + peer producer lane + the epilogue warps of both CTAs
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_holder, Cfg::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();          // barriers of both CTAs initialised, TMEM of both allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  vq_pdl_wait();
  const bool dyn = !GROUPED && p.sched != nullptr;
  // next work item of a consumer role (anything but the leader's producer lane): read the ring slot, release it on the leader
  auto next_work = [&](int& it) -> int {
    const int slot = it % SCHED_DEPTH;
    mbar_wait_cluster(&sfull_bar[slot], (it / SCHED_DEPTH) & 1);
    const int w = (int)sched_w[slot];
    mbar_arrive_cluster(&sempty_bar[slot], 0);
    ++it;
    return w;
  };

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      // leader: take the next item from the global counter and publish it to both CTAs; peer: take it from the ring. The
      // atomic for item i+1 is issued BEFORE the TMA loop of item i (its ~1 us round trip to L2 would otherwise sit between
      // two tiles of the producer); its result is first touched when item i+1 is published.
      // The first item of every pair is its static one (no atomic round trip before the kernel's first TMA load); the
      // counter hands out items npairs, npairs + 1, ...
      int w_ahead = pair;
      auto fetch = [&]() -> int {
        if (rank != 0) return next_work(it);
        const int slot = it % SCHED_DEPTH;
        mbar_wait_cluster(&sempty_bar[slot], ((it / SCHED_DEPTH) & 1) ^ 1);
        const int w = w_ahead;
        sched_w[slot] = (uint32_t)w;
        st_shared_cluster_u32(&sched_w[slot], 1, (uint32_t)w);
        mbar_arrive(&sfull_bar[slot]);
        mbar_arrive_release_cluster(&sfull_bar[slot], 1);
        ++it;
        if (w < total_work) w_ahead = npairs + (int)atomicAdd(p.sched, 1u);
        return w;
      };
      for (int wi = dyn ? fetch() : pair; wi < total_work; wi = dyn ? fetch() : wi + npairs)
      for (int w = wi * run; w < wi * run + run; ++w) {
        int g, split, m_blk, n_blk;
        decode(w, g, split, m_blk, n_blk);
        const CUtensorMap& tmA = tmA_[g];
        const CUtensorMap& tmB = tmB_[g];
        const int m0 = m_blk * 2 * GEMM_BM + (int)rank * GEMM_BM;        // this CTA's 128 rows of the 256-row tile
        const int n0 = n_blk * BN + (int)rank * (BN / 2);                // this CTA's half of the B tile
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kblocks, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);   // bytes of BOTH CTAs land on the leader's barrier
          if (!A_MN) {
            tma_load_2d_2sm(sa, &tmA, &full_bar[stage], kb * GEMM_BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < GEMM_BM / 64; ++j)
              tma_load_2d_2sm(sa + j * (GEMM_BK * 128), &tmA, &full_bar[stage], m0 + j * 64, kb * GEMM_BK);
          }
          if (!B_MN) {
            tma_load_2d_2sm(sb, &tmB, &full_bar[stage], kb * GEMM_BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < (BN / 2) / 64; ++j)
              tma_load_2d_2sm(sb + j * (GEMM_BK * 128), &tmB, &full_bar[stage], n0 + j * 64, kb * GEMM_BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      // dynamic schedule: this pair has taken its terminal index; the last pair to get here resets the counters for the next
      // launch that uses the slot (done here, under the last tile's MMAs and epilogue, not on the kernel's tail)
      if (dyn && rank == 0) {
        const uint32_t done = atomicAdd(p.sched + 1, 1u);
        if (done == (uint32_t)npairs - 1) {
          p.sched[0] = 0u;
          p.sched[1] = 0u;
          __threadfence();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) --------------------------------
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * GEMM_BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int astage = 0;
      uint32_t aphase = 0;
      int it = 0;
      for (int wi = dyn ? next_work(it) : pair; wi < total_work; wi = dyn ? next_work(it) : wi + npairs)
      for (int w = wi * run; w < wi * run + run; ++w) {
        const int split = GROUPED ? 0 : w / tiles_mn;
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
            const uint64_t adesc = A_MN ? umma_smem_desc_sw128(sa + k * 2048, GEMM_BK * 128, 1024)
                                        : umma_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? umma_smem_desc_sw128(sb + k * 2048, GEMM_BK * 128, 1024)
                                        : umma_smem_desc_sw128(sb + k * 32, 16, 1024);
            umma_f16_2sm(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage]);   // frees this stage in BOTH CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tfull_bar[astage]);    // accumulator complete -> epilogue warps of both CTAs
        if (++astage == 2) { astage = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue (both CTAs, own 128 rows) ----------------------------------
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const uint32_t stg = smem_u32(epi_stage) + (warp - 4) * EPI_TILE_BYTES;
    int astage = 0;
    uint32_t aphase = 0;
    int it = 0;
    // the ring slot is read by lane 0 of each epilogue warp (one arrival per warp on the leader's barrier) and broadcast
    auto warp_next = [&]() -> int {
      int w = 0;
      if (lane == 0) w = next_work(it);
      return __shfl_sync(0xffffffffu, w, 0);
    };
    for (int wi = dyn ? warp_next() : pair; wi < total_work; wi = dyn ? warp_next() : wi + npairs) {
    for (int w = wi * run; w < wi * run + run; ++w) {
      int g, split, m_blk, n_blk;
      decode(w, g, split, m_blk, n_blk);
      const bool has_k = split * kb_per_split < kblocks;
      const int row_base = m_blk * 2 * GEMM_BM + (int)rank * GEMM_BM + q * 32;
      const uint32_t t_base = tmem_base + astage * BN + ((uint32_t)(q * 32) << 16);
      uint64_t* tf = &tfull_bar[astage];
      if constexpr (GROUPED) {
        GemmArgs pg = p;                 // the epilogue is inlined: only the fields it reads exist
        pg.M = grp->M[g]; pg.N = grp->N[g]; pg.C = grp->C[g]; pg.ldc = grp->ldc[g];
        if (p.epi == EPI_ATOMIC_F32) gemm_epilogue_tile<EPI_ATOMIC_F32, BN>(pg, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase);
        else gemm_epilogue_tile<EPI_F32, BN>(pg, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase);
      } else
      if (p.epi == EPI_F32 && p.split_stride) {
        GemmArgs pg = p;
        pg.C = reinterpret_cast<float*>(p.C) + (size_t)split * p.split_stride;
        gemm_epilogue_tile<EPI_F32, BN>(pg, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase);
      } else
      switch (p.epi) {
        case EPI_BF16: gemm_epilogue_tile<EPI_BF16, BN>(p, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase); break;
        case EPI_RELU_BF16: gemm_epilogue_tile<EPI_RELU_BF16, BN>(p, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase); break;
        case EPI_RESID_F32: gemm_epilogue_tile<EPI_RESID_F32, BN>(p, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase); break;
        case EPI_ATOMIC_F32: gemm_epilogue_tile<EPI_ATOMIC_F32, BN>(p, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase); break;
        case EPI_RELUBWD_BF16: gemm_epilogue_tile<EPI_RELUBWD_BF16, BN>(p, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase); break;
        case EPI_ARGMAX: gemm_epilogue_tile<EPI_ARGMAX, BN>(p, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase); break;
        default: gemm_epilogue_tile<EPI_F32, BN>(p, t_base, stg, row_base, n_blk * BN, half, lane, has_k, tf, aphase); break;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty_bar[astage], 0);   // the leader's MMA lane owns accumulator reuse
      if (++astage == 2) { astage = 0; aphase ^= 1; }
    }
    if (!GROUPED && p.tail) {
      // ---- row tail: this CTA's 128 rows of block wi are complete in C (written by these 8 warps) and still in L2
      asm volatile("bar.sync 1, %0;" ::"n"(GEMM_EPI_WARPS * 32) : "memory");
      const int row0 = wi * 2 * GEMM_BM + (int)rank * GEMM_BM;
      if (p.tail == 1) gemm_tail_rmsnorm(p, row0, warp - 4, lane);
      else gemm_tail_visual(p, row0, warp - 4, lane);
    }
    }
  }

  tc_fence_before();
  cluster_sync_all();          // the peer's smem / TMEM must stay alive until the leader's last MMA and both epilogues are done
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

template <bool A_MN, bool B_MN>
__global__ void __maxnreg__(GEMM_MAXREG)
gemm_bf16_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs p) {
  gemm_pair_body<A_MN, B_MN, false>(&tmA, &tmB, p, nullptr);
}
// grouped form: GemmArgs carries K, the epilogue and alpha; M / N / C / ldc come from the group table
template <bool A_MN, bool B_MN>
__global__ void __maxnreg__(GEMM_MAXREG)
gemm_bf16_tcgen05_2cta_grouped_kernel(const __grid_constant__ GemmGroupMaps maps, const GemmArgs p, const __grid_constant__ GemmGroup grp) {
  gemm_pair_body<A_MN, B_MN, true>(maps.a, maps.b, p, &grp);
}

}  // namespace vq
