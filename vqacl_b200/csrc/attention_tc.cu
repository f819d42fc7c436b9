// Encoder self-attention on the 5th-generation tensor cores: tcgen05.mma with the score / output accumulators in TMEM,
// Q / K / V head slices staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) straight from the packed [rows, 3d] QKV
// activation, softmax done by threads that own ONE ROW each (tcgen05.ld 32x32b: thread = row), so the row maximum, the
// sum, the mask / bias adds and the dropout decisions need no shuffles at all.
//
// Replaces, for the encoder shape (Sq = Sk = S <= 64, relative bias on the text x text corner, additive key-padding mask —
// JointEncoder.forward, modeling_t5_our.py:225-301 + HF T5Attention, hf5.5 modeling_t5.py:277-338), the mma.sync kernel of
// attention.cu, whose 1100 instructions per warp around 64 HMMA made it issue-bound at 2.4x the HBM time
// (profiles/r01_step_kernel_metrics.md).
//
// One work item = one batch element x TWO adjacent heads, stacked into one M = 128 MMA:
//   S2[128 x 128] = Q2[128 x 64] K2[128 x 64]^T   rows 0-63: head h queries, rows 64-127: head h+1; columns likewise for keys.
//                                                 Only the two diagonal 64 x 64 blocks are meaningful; the MMA is ~1 % of the
//                                                 step's FLOPs, so computing the off-diagonal blocks costs nothing.
//   softmax rows -> P2[128 x 128] bf16 in shared memory, block diagonal (the off-diagonal halves are zeroed once and never
//                                                 written), in the K-major 128-byte-swizzle operand layout
//   O2[128 x 64]  = P2[128 x 128] V2[128 x 64]    V2 = the two heads' V slices stacked along the key axis (MN-major operand)
// Roles (forward): warps 0-7 softmax + epilogue, TWO threads per stacked query row (TMEM lane = row; each thread holds 32 of
// the row's 64 scores), warp 8 per-item header + TMA producer + MMA issuer + TMEM owner; two shared-memory stages, two CTAs per
// SM; the S MMA of item i+1 runs under the softmax of item i and the O epilogue of item i-1 is folded into item i's softmax
// phase, so the softmax warps never wait for an MMA. O rows leave as 8 rows x 64 contiguous bytes per store instruction,
// transposed through shared-memory slots the thread owns anyway. The backward kernel (further down) has its own role layout.
#include "gemm.h"
#include "ops.h"

#include <stdlib.h>

namespace vq {

constexpr int AT_S_TC = 64;                          // positions per head tile
constexpr int TC_P_BYTES = 2 * 16384;               // two K-atoms (keys of head h | keys of head h+1), each 128 rows x 128 B
constexpr int TC_TMEM_COLS = 512;

// ---- forward: small CTAs (2 x 48 KB of shared memory, 128 TMEM columns, <= 128 registers) so that TWO are resident per SM;
//      inside a CTA the compute phases of one item run back to back (S -> softmax -> O -> store) under the TMA loads of the
//      next item, and the other CTA of the SM covers the hand-over latencies. History (profiles/r02_attention_tc.md):
//      v1, one CTA per SM with two stages, two S accumulators and a deferred epilogue: correct, 58 us (mma.sync kernel: 40 us);
//      ncu: 6 warps per SM, issue slots 16 % busy. v2, three single-stage CTAs per SM: 55 us; ncu: 37 % of the stall samples
//      are the row warps waiting for the item's TMA loads (no prefetch), DRAM 22 %. v3, v2 + a second stage, two CTAs per
//      SM: 58 us; the loads are hidden but 8 row warps per SM issue one instruction per 11 cycles each (thread = row keeps 64
//      scores per thread and ~1200 dependent instructions per item). v4 (this one): two threads per row -> 16 row warps per
//      SM, 32 scores per thread in registers, a single pass over TMEM; the halves of the row maximum / sum are exchanged
//      through shared memory with a 64-thread named barrier per warp pair.
constexpr int TCF_STAGE_BYTES = 3 * 16384;          // Q2 | K2 | V2; the P tile overwrites Q2 | K2 once the S MMA has completed
constexpr int TCF_HDR_FLOATS = 64 + 2 * 128;        // kmask[64] | bias[2][128]
constexpr int TCF_STAGES = 2;                       // the TMA loads of item n+1 are in flight while item n is computed
constexpr int TCF_XCHG_FLOATS = 256 + 2 * 512;      // row maximum halves [2][128] | row sum halves, one copy per item parity, [2][128] each at +512
constexpr int TCF_SMEM_BYTES = 1024 + TCF_STAGES * TCF_STAGE_BYTES + 2 * TCF_HDR_FLOATS * 4 + TCF_XCHG_FLOATS * 4 + 128;
constexpr int TCF_ROW_WARPS = 8;                    // warps 0-7: TWO threads per stacked query row (32 of its 64 score columns each)
constexpr int TCF_THREADS = (TCF_ROW_WARPS + 1) * 32;   // + warp 8: header, TMA, MMA issue, TMEM
constexpr int TCF_TMEM_COLS = 256;                  // S2 [0,128) | O2 [128,192)
constexpr int TCF_CTAS_PER_SM = 2;

struct AttnTcArgs {
  __nv_bfloat16* o; int ldo;
  float* lse;
  int B, H, S, Lt;
  const float* rel_table;       // [buckets, H]
  const float* keymask;         // [B, S] additive
  uint32_t drop_thr; float drop_inv_keep; uint32_t seed;
};

VQ_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
VQ_DEVINL uint4 lds128_u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
VQ_DEVINL void sts128_u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(TCF_THREADS, TCF_CTAS_PER_SM)
attn_enc_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const AttnTcArgs p, const __grid_constant__ AttnBuckets bk) {
  vq_pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_u32 + 1023u) & ~1023u) - raw_u32);
  float* hdr = reinterpret_cast<float*>(smem + TCF_STAGES * TCF_STAGE_BYTES);
  float* xchg = hdr + 2 * TCF_HDR_FLOATS;      // [2 halves][128 rows] maxima | [2][128] sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + TCF_XCHG_FLOATS);
  uint64_t* full_bar = bars;        // [2] TMA bytes of a stage landed (+ header written)
  uint64_t* sfull_bar = bars + 2;   // S accumulator complete
  uint64_t* sread_bar = bars + 3;   // S read by the row warps (=> the next item's S MMA may overwrite it)
  uint64_t* pfull_bar = bars + 4;   // P tile written (and the previous item's O read) by the row warps
  uint64_t* ofull_bar = bars + 5;   // O accumulator complete (=> the stage's shared memory is free)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hp = p.H >> 1;
  const int nitems = p.B * hp;

  if (warp == TCF_ROW_WARPS) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      mbar_init(&full_bar[0], 1);
      mbar_init(&full_bar[1], 1);
      mbar_init(sfull_bar, 1);
      mbar_init(sread_bar, TCF_ROW_WARPS);
      mbar_init(pfull_bar, TCF_ROW_WARPS);
      mbar_init(ofull_bar, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_holder, TCF_TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  vq_pdl_wait();

  if (warp == TCF_ROW_WARPS) {
    // ------------------------------------------------ header + TMA + MMA issue ------------------------------------------------
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, false, true);
    // Per-item header (key mask of the batch element, relative-position bias of the two heads): the global loads are issued
    // early into registers (hdr_fetch), the shared-memory copy is written — and the TMA loads of the item issued — only once
    // the stage is free (load_item), so that no global-memory latency sits between "stage free" and "TMA in flight".
    float hk[2], hb[8];
    auto hdr_fetch = [&](int it) {
      const int b = it / hp, h = (it - b * hp) * 2;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = lane + 32 * u;
        hk[u] = j < p.S ? (p.keymask ? p.keymask[(size_t)b * p.S + j] : 0.f) : -INFINITY;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = lane + 32 * u, hd = r >> 7, rel = r & 127;
        hb[u] = rel < 127 ? p.rel_table[(int)bk.b[rel] * p.H + h + hd] : 0.f;
      }
    };
    auto load_item = [&](int m, int it) {
      const int b = it / hp, h = (it - b * hp) * 2;
      float* hs = hdr + (m & 1) * TCF_HDR_FLOATS;
#pragma unroll
      for (int u = 0; u < 2; ++u) hs[lane + 32 * u] = hk[u];
#pragma unroll
      for (int u = 0; u < 8; ++u) hs[64 + lane + 32 * u] = hb[u];
      __syncwarp();
      if (lane == 0) {
        uint8_t* st = smem + (m & 1) * TCF_STAGE_BYTES;
        uint64_t* fb = &full_bar[m & 1];
        mbar_expect_tx(fb, TCF_STAGE_BYTES);
        const int row = b * p.S;
        tma_load_2d(st, &tmQ, fb, h * 64, row);
        tma_load_2d(st + 8192, &tmQ, fb, (h + 1) * 64, row);
        tma_load_2d(st + 16384, &tmK, fb, h * 64, row);
        tma_load_2d(st + 16384 + 8192, &tmK, fb, (h + 1) * 64, row);
        tma_load_2d(st + 32768, &tmV, fb, h * 64, row);
        tma_load_2d(st + 32768 + 8192, &tmV, fb, (h + 1) * 64, row);
      }
    };
    // S2 = Q2 K2^T of item m (stage m & 1); the caller has made sure the rows have read the previous S2
    auto issue_s = [&](int m) {
      const uint32_t sq = smem_u32(smem + (m & 1) * TCF_STAGE_BYTES), sk = sq + 16384;
      mbar_wait(&full_bar[m & 1], (m >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16(tmem_base, umma_smem_desc_sw128(sq + k * 32, 16, 1024), umma_smem_desc_sw128(sk + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
      umma_commit(sfull_bar);
    };
    const int it0 = blockIdx.x, gs = gridDim.x;
    if (it0 < nitems) { hdr_fetch(it0); load_item(0, it0); }
    if (it0 + gs < nitems) { hdr_fetch(it0 + gs); load_item(1, it0 + gs); }
    if (it0 < nitems && lane == 0) issue_s(0);
    __syncwarp();
    int n = 0;
    for (int it = it0; it < nitems; it += gs, ++n) {
      if (it + 2 * gs < nitems) hdr_fetch(it + 2 * gs);      // in flight while lane 0 waits below
      if (lane == 0) {
        const uint32_t sq = smem_u32(smem + (n & 1) * TCF_STAGE_BYTES), sv = sq + 32768;
        if (it + gs < nitems) {
          mbar_wait(sread_bar, n & 1);                  // the rows hold S2 of item n in registers: the next S MMA runs under their softmax
          issue_s(n + 1);
        }
        mbar_wait(pfull_bar, n & 1);                    // P tile of item n written; O2 of item n-1 read (the rows do that first)
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_f16(tmem_base + 128, umma_smem_desc_sw128(sq + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                   umma_smem_desc_sw128(sv + kk * 2048, 8192, 1024), idesc_o, kk > 0 ? 1u : 0u);
        umma_commit(ofull_bar);
        if (it + 2 * gs < nitems) mbar_wait(ofull_bar, n & 1);   // O MMA done: stage n & 1 (V2, P tile) and header copy n & 1 are free
      }
      __syncwarp();
      if (it + 2 * gs < nitems) load_item(n + 2, it + 2 * gs);
    }
  } else {
    // ------------------------------------------------ softmax + epilogue (two threads per stacked row) ------------------------------------------------
    const int rw = warp & 3, ch = warp >> 2;   // TMEM lane quarter (a warp may only touch lanes 32 * (warp % 4) ...) | column half
    const int r = rw * 32 + lane;              // stacked query row = TMEM lane
    const int hsel = r >> 6, q = r & 63;
    const uint32_t lane_addr = (uint32_t)(rw * 32) << 16;
    const uint32_t sP_row0 = smem_u32(smem) + hsel * 16384 + r * 128;         // this row's 128 B of the key atom of its own head
    const uint32_t sZ_row0 = smem_u32(smem) + (hsel ^ 1) * 16384 + r * 128;   // ... and of the other head's atom (zeros: block diagonal)
    const int sw = r & 7;
    const uint32_t ts = tmem_base + hsel * 64 + ch * 32 + lane_addr;          // this thread's 32 score columns
    float* xmax = xchg + r, *xsum = xchg + 256 + r;                            // [half][row]
    const int bar_id = 1 + rw;                                                 // named barrier of the warp pair (rw, rw + 4)
    // O2 row of the item with running index m (its row sum halves are in xsum, parity m & 1): scale, convert, store
    float m_prev = 0.f;
    int b_prev = 0, h_prev = 0;
    // A thread-per-row store is 32 row-strided 16-byte requests per instruction (measured on the backward kernel: 2.5 us of a
    // 6.7 us item). The 64 bytes of each row are therefore transposed through shared memory and leave as 8 rows x 64 contiguous
    // bytes per instruction. The staging needs no memory of its own: a thread parks its row in the four 16-byte slots of the
    // OTHER head's key atom that it zeroes for the block-diagonal P tile anyway (stage `st_stage`: operands consumed, P tile not
    // yet written) — slots only this warp ever touches, read back by the warp's own lanes after a __syncwarp.
    const int rr0 = lane >> 2, piece = lane & 3;
    const uint32_t zslot_rd0 = smem_u32(smem) + (hsel ^ 1) * 16384 + (rw * 32 + rr0) * 128 + (((ch * 4 + piece) ^ rr0) << 4);
    const int qb = (rw & 1) * 32;
    auto epilogue = [&](int m, int st_stage) {
      mbar_wait(ofull_bar, m & 1);
      tc_fence_after();
      const float l = xsum[(m & 1) * 512] + xsum[(m & 1) * 512 + 128];
      if (ch == 0 && q < p.S && p.lse) p.lse[((size_t)b_prev * p.H + h_prev + hsel) * p.S + q] = m_prev + __logf(l);
      const float osc = (p.drop_thr ? p.drop_inv_keep : 1.f) / l;
      uint32_t o0[32];
      tmem_ld_32x32(tmem_base + 128 + lane_addr + ch * 32, o0);
      tmem_ld_wait();
      tc_fence_before();
      const uint32_t zw = sZ_row0 + st_stage * TCF_STAGE_BYTES;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        sts128_u(zw + ((uint32_t)((ch * 4 + c) ^ sw) << 4),
                 pack_bf16(__uint_as_float(o0[8 * c]) * osc, __uint_as_float(o0[8 * c + 1]) * osc),
                 pack_bf16(__uint_as_float(o0[8 * c + 2]) * osc, __uint_as_float(o0[8 * c + 3]) * osc),
                 pack_bf16(__uint_as_float(o0[8 * c + 4]) * osc, __uint_as_float(o0[8 * c + 5]) * osc),
                 pack_bf16(__uint_as_float(o0[8 * c + 6]) * osc, __uint_as_float(o0[8 * c + 7]) * osc));
      __syncwarp();
      char* gp = reinterpret_cast<char*>(p.o + ((size_t)b_prev * p.S + qb + rr0) * p.ldo + (h_prev + hsel) * 64 + ch * 32 + piece * 8);
      const size_t gstep = (size_t)8 * p.ldo * 2;
      const uint32_t zr = zslot_rd0 + st_stage * TCF_STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 u = lds128_u(zr + i * 1024);
        if (qb + rr0 + 8 * i < p.S) *reinterpret_cast<uint4*>(gp + i * gstep) = u;
      }
      __syncwarp();
    };
    int n = 0;
    for (int it = blockIdx.x; it < nitems; it += gridDim.x, ++n) {
      const int b = it / hp, h = (it - b * hp) * 2;
      const float* hs = hdr + (n & 1) * TCF_HDR_FLOATS;
      const float* bp = hs + 64 + hsel * 128 + (AT_S_TC - 1) - q;
      const bool biased = ch == 0 && q < p.Lt;       // the biased text x text corner lies in columns < Lt <= 32
      const uint32_t sP_row = sP_row0 + (n & 1) * TCF_STAGE_BYTES, sZ_row = sZ_row0 + (n & 1) * TCF_STAGE_BYTES;
      mbar_wait(&full_bar[n & 1], (n >> 1) & 1);   // header visible
      mbar_wait(sfull_bar, n & 1);
      tc_fence_after();
      float x[32];
      {
        uint32_t s0[32];
        tmem_ld_32x32(ts, s0);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sread_bar);       // S2 is in registers: the next item's S MMA may start
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 km = *reinterpret_cast<const float4*>(hs + ch * 32 + 4 * j4);
          x[4 * j4] = __uint_as_float(s0[4 * j4]) + km.x;
          x[4 * j4 + 1] = __uint_as_float(s0[4 * j4 + 1]) + km.y;
          x[4 * j4 + 2] = __uint_as_float(s0[4 * j4 + 2]) + km.z;
          x[4 * j4 + 3] = __uint_as_float(s0[4 * j4 + 3]) + km.w;
        }
      }
      if (biased) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < p.Lt) x[j] += bp[j];
      }
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < 32; ++j) mx[j & 3] = fmaxf(mx[j & 3], x[j]);
      float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      xmax[ch * 128] = m;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      m = fmaxf(m, xmax[(ch ^ 1) * 128]);
      // e = exp(x - m), un-normalised: 1 / sum and the dropout scale are applied to the O row
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[16];
      const uint32_t pi0 = attn_pair_idx((uint32_t)(b * p.H + h + hsel), q, ch * 32);
#pragma unroll
      for (int j2 = 0; j2 < 16; ++j2) {
        float e0 = __expf(x[2 * j2] - m), e1 = __expf(x[2 * j2 + 1] - m);
        l4[j2 & 3] += e0 + e1;
        if (p.drop_thr) {
          const uint32_t hh = vq_hash_pair(p.seed, pi0 + j2);
          e0 = (hh & 0xFFFFu) >= p.drop_thr ? e0 : 0.f;
          e1 = (hh >> 16) >= p.drop_thr ? e1 : 0.f;
        }
        pk[j2] = pack_bf16(e0, e1);
      }
      // the previous item's O row: its MMA finished long ago (nobody waits here), and reading it now frees the O accumulator
      // before this item's P tile is released to the MMA warp
      if (n > 0) epilogue(n - 1, n & 1);
      xsum[(n & 1) * 512 + ch * 128] = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      m_prev = m; b_prev = b; h_prev = h;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t off = (uint32_t)((ch * 4 + c) ^ sw) << 4;
        sts128_u(sP_row + off, pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        sts128_u(sZ_row + off, 0u, 0u, 0u, 0u);
      }
      fence_proxy_async();
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");     // the partner's half of the row sum is visible to the next epilogue
      if (lane == 0) mbar_arrive(pfull_bar);
    }
    if (n > 0) epilogue(n - 1, (n - 1) & 1);       // its own stage: the O MMA the epilogue waits for was the last reader
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TCF_ROW_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TCF_TMEM_COLS);
  }
}

// =====================================================================================================================
// Backward of the same problem on the same machinery. Per item (batch element x two stacked heads):
//   S2  = Q2 K2^T, dP2 = dO2 V2^T                       (two M = 128, N = 128 MMAs into TMEM)
//   row threads: P = exp(S + mask + bias - lse), dP *= dropout, D = sum_j P dP (row-local: no shuffles),
//                dS = P (dP - D); P (dropout folded in) and dS go to two block-diagonal bf16 tiles in shared memory
//   dV2 = P2^T dO2, dK2 = dS2^T Q2   (the tiles read as MN-major A operands), dQ2 = dS2 K2   (read as K-major A operand)
//   row threads: TMEM -> bf16 -> full 128-byte rows of dQ / dK / dV
// The bias-table gradient (sums of dS along the diagonals of the text x text corner) is accumulated in a per-CTA shared
// table across all items and flushed with one global atomic per (bucket, head) at the end of the kernel.
// =====================================================================================================================
// -DVQ_ATTN_TRACE (VQACL_NVCC_EXTRA of build.py): CTA 0 stamps clock64() at every hand-off of its first 16 items into a
// buffer installed with vqacl_debug_attn_trace — tools/attn_trace.py turns the stamps into a per-item timeline.
#ifdef VQ_ATTN_TRACE
__device__ long long* g_attn_trace = nullptr;     // [role 0..3][item 0..15][slot 0..7]
#define VQ_TR(role, n, slot)                                                                                  \
  do {                                                                                                        \
    if (blockIdx.x == 0 && (n) < 16 && g_attn_trace) g_attn_trace[((role) * 16 + (n)) * 8 + (slot)] = clock64(); \
  } while (0)
#else
#define VQ_TR(role, n, slot) do { } while (0)
#endif

constexpr int TCB_STAGE_BYTES = 4 * 16384;          // Q2 | K2 | V2 | dO2
constexpr int TCB_HDR_FLOATS = 64 + 2 * 128 + 128;  // kmask[64] | bias[2][128] | lse[128] (stacked rows; +inf beyond S)
constexpr int TCB_DTAB_FLOATS = 32 * 16;            // [bucket][head] accumulation table (32 buckets, H <= 16)
constexpr int TCB_DX_FLOATS = 2 * 4 * 128;          // row-sum quarters exchanged between the 4 threads of a row, per item parity
constexpr int TCB_ROW_WARPS = 16;                   // FOUR threads per stacked query row (16 of its 64 columns each)
constexpr int TCB_THREADS = (TCB_ROW_WARPS + 2) * 32;   // warp 0: header + TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-17: rows.
                                                        // The two single-thread warps come FIRST: as the youngest warps of their
                                                        // schedulers (warps 16 / 17) they were starved by the row warps' spin loops —
                                                        // 2 us from "stage free" to "TMA issued", 50 ns per tcgen05.mma issued.
constexpr int TCB_PRODUCER_WARP = 0, TCB_ISSUER_WARP = 1, TCB_FIRST_ROW_WARP = 2;
constexpr int TCB_OUT_TILE = 2048;                  // output staging: 32 rows x 64 bytes per epilogue warp (12 of the 16 row warps)
constexpr int TCB_SMEM_BYTES = 1024 + 2 * TCB_STAGE_BYTES + 2 * TC_P_BYTES + (2 * TCB_HDR_FLOATS + TCB_DTAB_FLOATS + TCB_DX_FLOATS) * 4 + 12 * TCB_OUT_TILE + 256;
static_assert(TCB_SMEM_BYTES <= 232448, "attention backward: shared memory");

struct AttnTcBwdArgs {
  __nv_bfloat16 *dq, *dk, *dv; int lddq, lddk, lddv;
  const float* lse;
  int B, H, S, Lt;
  const float* rel_table; const float* keymask;
  float* d_rel_table;
  uint32_t drop_thr; float drop_inv_keep; uint32_t seed;
};

// Structure (history with timings and traces: profiles/r02_attention_tc.md; 99.5 us -> 69.9 us, mma.sync kernel: 91 us).
// One CTA per SM (two 64 KB stages + the P and dS tiles), so the warps come from FOUR threads per row: 16 row warps, each
// thread holds 16 scores and 16 dP values, the row's D = sum_j P dP is exchanged between its four threads through shared
// memory (128-thread named barrier per TMEM lane quarter). S and dP of item n+1 are issued as soon as the rows hold item n's
// in registers and its operands have landed; the gradient MMAs of item n as soon as its P / dS tiles are written — one issuer
// thread polls both conditions. The dQ / dK / dV rows of item n are stored while item n+1 is being computed (deferred
// epilogue), through per-warp swizzled staging tiles so that a store instruction covers 8 rows x 64 contiguous bytes; the
// four warps without an output sum the bias-gradient diagonals meanwhile. The per-item header (key mask, bias, log-sum-exp)
// is prefetched into registers an item early and written to shared memory when the rows release the slot, so that only the
// TMA issue itself sits between "stage free" and "loads in flight".
__global__ void __launch_bounds__(TCB_THREADS, 1)
attn_enc_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, const AttnTcBwdArgs p,
                       const __grid_constant__ AttnBuckets bk) {
  vq_pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_u32 + 1023u) & ~1023u) - raw_u32);
  uint8_t* sP = smem + 2 * TCB_STAGE_BYTES;
  uint8_t* sdS = sP + TC_P_BYTES;
  float* hdr = reinterpret_cast<float*>(sdS + TC_P_BYTES);
  float* dtab = hdr + 2 * TCB_HDR_FLOATS;
  float* dx = dtab + TCB_DTAB_FLOATS;
  uint8_t* ostage = reinterpret_cast<uint8_t*>(dx + TCB_DX_FLOATS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ostage + 12 * TCB_OUT_TILE);
  uint64_t* full_bar = bars;            // [2] TMA bytes landed (+ header written)
  uint64_t* empty_bar = bars + 2;       // [2] the item's last MMAs are complete: its stage is free
  uint64_t* sfull_bar = bars + 4;       // S and dP accumulators complete
  uint64_t* sread_bar = bars + 5;       // ... held in registers by the 16 row warps
  uint64_t* pfull_bar = bars + 6;       // [2, by stage] P and dS tiles written (and the previous item's outputs read): also "the rows are
                                        // done with this stage's header", which is what the producer waits for before rewriting it
  uint64_t* ofull_bar = bars + 8;       // dQ, dK, dV accumulators complete
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hp = p.H >> 1;
  const int nitems = p.B * hp;
  const int it0 = blockIdx.x, gs = gridDim.x;

  if (warp == TCB_PRODUCER_WARP && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(sfull_bar, 1);
    mbar_init(sread_bar, TCB_ROW_WARPS);
    mbar_init(&pfull_bar[0], TCB_ROW_WARPS);
    mbar_init(&pfull_bar[1], TCB_ROW_WARPS);
    mbar_init(ofull_bar, 1);
    mbar_fence_init();
  }
  if (warp == TCB_ISSUER_WARP) tmem_alloc(tmem_holder, TC_TMEM_COLS);
  for (int i = threadIdx.x; i < 2 * TC_P_BYTES / 16; i += TCB_THREADS) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0, 0, 0, 0);   // P and dS tiles
  for (int i = threadIdx.x; i < TCB_DTAB_FLOATS; i += TCB_THREADS) dtab[i] = 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // TMEM columns: S [0,128) | dP [128,256) | dQ [256,320) | dK [320,384) | dV [384,448)
  vq_pdl_wait();

  if (warp == TCB_PRODUCER_WARP) {
    // ------------------------------------------------ header + TMA producer ------------------------------------------------
    float hk[2], hb[8], hl[4];
    auto hdr_fetch = [&](int it) {
      const int b = it / hp, h = (it - b * hp) * 2;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = lane + 32 * u;
        hk[u] = j < p.S ? (p.keymask ? p.keymask[(size_t)b * p.S + j] : 0.f) : -INFINITY;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = lane + 32 * u, hd = r >> 7, rel = r & 127;
        hb[u] = rel < 127 ? p.rel_table[(int)bk.b[rel] * p.H + h + hd] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = lane + 32 * u, hd = r >> 6, q = r & 63;
        hl[u] = q < p.S ? p.lse[((size_t)b * p.H + h + hd) * p.S + q] : INFINITY;     // rows >= S: P = dS = 0
      }
    };
    int n = 0;
    if (it0 < nitems) hdr_fetch(it0);
    for (int it = it0; it < nitems; it += gs, ++n) {
      const int stage = n & 1;
      const int b = it / hp, h = (it - b * hp) * 2;
      // Header first, as soon as the ROWS are done with this stage's previous header (their pfull arrival — the gradient MMAs of
      // that item, whose completion frees the operand half of the stage, are only being issued now): measured, the 14 STS of this
      // warp take 1.6-2.4 us under MIO throttle while the tensor pipe and 16 row warps are busy, and they used to sit between
      // "stage free" and "TMA issued", i.e. on the loop that sets the period of the whole pipeline.
      if (lane == 0) { VQ_TR(0, n, 0); if (n >= 2) mbar_wait(&pfull_bar[stage], ((n >> 1) & 1) ^ 1); }
      __syncwarp();
      float* hs = hdr + stage * TCB_HDR_FLOATS;
#pragma unroll
      for (int u = 0; u < 2; ++u) hs[lane + 32 * u] = hk[u];
#pragma unroll
      for (int u = 0; u < 8; ++u) hs[64 + lane + 32 * u] = hb[u];
#pragma unroll
      for (int u = 0; u < 4; ++u) hs[64 + 256 + lane + 32 * u] = hl[u];
      __syncwarp();
      if (lane == 0) VQ_TR(0, n, 3);
      if (it + gs < nitems) hdr_fetch(it + gs);      // the next item's header: in flight across the wait below and the whole next item
      if (lane == 0) { mbar_wait(&empty_bar[stage], ((n >> 1) & 1) ^ 1); VQ_TR(0, n, 1); }
      __syncwarp();
      if (lane == 0) {
        uint8_t* st = smem + stage * TCB_STAGE_BYTES;
        VQ_TR(0, n, 4);
        mbar_expect_tx(&full_bar[stage], TCB_STAGE_BYTES);
        const int row = b * p.S;
        tma_load_2d(st, &tmQ, &full_bar[stage], h * 64, row);
        tma_load_2d(st + 8192, &tmQ, &full_bar[stage], (h + 1) * 64, row);
        VQ_TR(0, n, 5);
        tma_load_2d(st + 16384, &tmK, &full_bar[stage], h * 64, row);
        tma_load_2d(st + 16384 + 8192, &tmK, &full_bar[stage], (h + 1) * 64, row);
        VQ_TR(0, n, 6);
        tma_load_2d(st + 32768, &tmV, &full_bar[stage], h * 64, row);
        tma_load_2d(st + 32768 + 8192, &tmV, &full_bar[stage], (h + 1) * 64, row);
        VQ_TR(0, n, 7);
        tma_load_2d(st + 49152, &tmdO, &full_bar[stage], h * 64, row);
        tma_load_2d(st + 49152 + 8192, &tmdO, &full_bar[stage], (h + 1) * 64, row);
        VQ_TR(0, n, 2);
      }
    }
  } else if (warp == TCB_ISSUER_WARP) {
    // ------------------------------------------------ MMA issuer ------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, false, false);     // S, dP: both operands K-major
      constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64, true, true);        // dV, dK: A = tile^T (MN-major), B MN-major
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, 64, false, true);       // dQ: A = dS tile K-major, B MN-major
      const uint32_t sP_u32 = smem_u32(sP), sdS_u32 = smem_u32(sdS);
      auto issue_sdp = [&](int m) {
        const int stage = m & 1;
        VQ_TR(1, m, 1);
        tc_fence_after();
        const uint32_t sq = smem_u32(smem + stage * TCB_STAGE_BYTES), sk = sq + 16384, sv = sq + 32768, sdo = sq + 49152;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base, umma_smem_desc_sw128(sq + k * 32, 16, 1024), umma_smem_desc_sw128(sk + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base + 128, umma_smem_desc_sw128(sdo + k * 32, 16, 1024), umma_smem_desc_sw128(sv + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(sfull_bar);
        VQ_TR(1, m, 2);
      };
      auto issue_grads = [&](int n) {
        const int stage = n & 1;
        tc_fence_after();
        const uint32_t sq = smem_u32(smem + stage * TCB_STAGE_BYTES), sk = sq + 16384, sdo = sq + 49152;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {      // contraction over the 128 stacked query rows, 16 per step
          const uint64_t a_p = umma_smem_desc_sw128(sP_u32 + kk * 2048, 16384, 1024);
          const uint64_t a_s = umma_smem_desc_sw128(sdS_u32 + kk * 2048, 16384, 1024);
          umma_f16(tmem_base + 384, a_p, umma_smem_desc_sw128(sdo + kk * 2048, 8192, 1024), idesc_t, kk > 0 ? 1u : 0u);   // dV = P^T dO
          umma_f16(tmem_base + 320, a_s, umma_smem_desc_sw128(sq + kk * 2048, 8192, 1024), idesc_t, kk > 0 ? 1u : 0u);    // dK = dS^T Q
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)         // contraction over the 128 stacked keys
          umma_f16(tmem_base + 256, umma_smem_desc_sw128(sdS_u32 + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                   umma_smem_desc_sw128(sk + kk * 2048, 8192, 1024), idesc_q, kk > 0 ? 1u : 0u);                          // dQ = dS K
        umma_commit(ofull_bar);
        umma_commit(&empty_bar[stage]);
        VQ_TR(1, n, 5);
      };
      // Two independent event streams share this thread: "S / dP of item m may be issued" (the rows hold item m-1 in registers
      // AND item m's operands have landed) and "dV / dK / dQ of item n may be issued" (its P / dS tiles are written). Issuing in
      // program order made the gradient MMAs of item n — whose completion frees the stage the item after next is loaded into —
      // wait for the TMA of item n+1 (measured: the period of the whole pipeline); poll both and issue whichever is ready.
      const int my_items = it0 < nitems ? (nitems - it0 + gs - 1) / gs : 0;
      int m = 0, n = 0;                                  // next S/dP item, next gradient item
      uint32_t spins = 0;
      while (n < my_items) {
        bool progressed = false;
        if (m < my_items && (m == 0 || mbar_try_wait(sread_bar, (m - 1) & 1)) && mbar_try_wait(&full_bar[m & 1], (m >> 1) & 1)) {
          issue_sdp(m);
          ++m;
          progressed = true;
        }
        if (n < m && mbar_try_wait(&pfull_bar[n & 1], (n >> 1) & 1)) {
          VQ_TR(1, n, 4);
          issue_grads(n);
          ++n;
          progressed = true;
        }
        if (progressed) spins = 0;
        else if (++spins > VQ_MBAR_SPIN_LIMIT) {
          printf("vqacl_b200: attention backward issuer timed out (block %d)\n", blockIdx.x);
          __trap();
        }
      }
    }
  } else {
    // ------------------------------------------------ row threads: 4 per stacked row ------------------------------------------------
    const int rw = warp & 3, cq = (warp - TCB_FIRST_ROW_WARP) >> 2;    // TMEM lane quarter (fixed by the warp id) | column quarter
    const int r = rw * 32 + lane;
    const int hsel = r >> 6, q = r & 63;
    const uint32_t lane_addr = (uint32_t)(rw * 32) << 16;
    const uint32_t row_off = hsel * 16384 + r * 128, zero_off = (hsel ^ 1) * 16384 + r * 128;
    const uint32_t sP_u = smem_u32(sP), sdS_u = smem_u32(sdS);
    const int sw = r & 7;
    const int bar_id = 1 + rw;                   // the four warps that share TMEM lane quarter rw
    float* dxr = dx + r;                         // [parity][column quarter][row]
    int b_prev = 0, h_prev = 0;
    const int trole = warp == TCB_FIRST_ROW_WARP ? 2 : 3;         // trace: first and last row warp
    const bool tron = (warp == TCB_FIRST_ROW_WARP || warp == TCB_FIRST_ROW_WARP + TCB_ROW_WARPS - 1) && lane == 0;
    (void)trole; (void)tron;
#define VQ_TRR(n, slot) do { if (tron) VQ_TR(trole, n, slot); } while (0)
    // Epilogue roles: the warps of column quarter 0 / 1 / 2 store dQ / dK / dV (thread = row, all 64 columns of the head, in two
    // halves of 32); column quarter 3 sums the bias-gradient diagonals instead. A thread-per-row store is 32 row-strided 16-byte
    // requests per instruction (3 072 per item: measured 2.5 us of the item's 6.7), so each half is transposed through a private
    // XOR-swizzled 32 x 64 B tile and leaves as 8 rows x 64 contiguous bytes per instruction.
    const int rr0 = lane >> 2, piece = lane & 3;                       // readback role: row inside an 8-row group, 16-byte piece
    const uint32_t stg = smem_u32(ostage) + (uint32_t)((cq < 3 ? cq * 4 + rw : 0) * TCB_OUT_TILE);
    const uint32_t sts_base = stg + lane * 64 + (((lane >> 1) & 3) << 4);
    const uint32_t lds_base = stg + rr0 * 64 + ((piece ^ ((rr0 >> 1) & 3)) << 4);
    const int qb = (rw & 1) * 32;                                       // first row (inside its head) of this warp's 32-row slab
    __nv_bfloat16* const outp = cq == 0 ? p.dq : (cq == 1 ? p.dk : p.dv);
    const int outld = cq == 0 ? p.lddq : (cq == 1 ? p.lddk : p.lddv);
    // Bias-table gradient of the item whose dS tile is in shared memory: sums along the diagonals of the text x text corner, one
    // thread per (head, diagonal), by the four warps (column quarter 3) that have no output rows to store — they do it while the
    // other twelve run the epilogue of the same item, so it is off the rows -> MMA critical path. Named barrier 5 = "every warp
    // has written its part of the dS tile" (arrive: quarters 0-2, sync: quarter 3), barrier 6 = "the diagonals are summed" (the
    // other way round; waited for before the tile is rewritten).
    auto diagonals = [&]() {
      asm volatile("bar.sync 5, %0;" ::"n"(TCB_ROW_WARPS * 32) : "memory");
      const int ndiag = 2 * p.Lt - 1;
      if (q < ndiag) {
        const int rel = q - (p.Lt - 1);             // key - query
        const int q_lo = max(0, -rel), q_hi = min(p.Lt, p.Lt - rel);
        float acc = 0.f;
        const uint8_t* tile = sdS + hsel * 16384;
#pragma unroll 4
        for (int qi = q_lo; qi < q_hi; ++qi) {
          const int row = hsel * 64 + qi, j = qi + rel;
          acc += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(tile + row * 128 + (((j >> 3) ^ (row & 7)) << 4) + (j & 7) * 2));
        }
        atomicAdd(&dtab[(int)bk.b[rel + (AT_S_TC - 1)] * 16 + h_prev + hsel], acc);
      }
    };
    auto epilogue = [&](int m) {
      mbar_wait(ofull_bar, m & 1);
      VQ_TRR(m + 1, 4);
      tc_fence_after();
      if (cq < 3) {
        char* gp = reinterpret_cast<char*>(outp + ((size_t)b_prev * p.S + qb + rr0) * outld + (h_prev + hsel) * 64 + piece * 8);
        const size_t gstep = (size_t)8 * outld * 2;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t o[32];
          tmem_ld_32x32(tmem_base + 256 + cq * 64 + half * 32 + lane_addr, o);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128_u(sts_base ^ (j << 4), pack_bf16(__uint_as_float(o[8 * j]), __uint_as_float(o[8 * j + 1])),
                     pack_bf16(__uint_as_float(o[8 * j + 2]), __uint_as_float(o[8 * j + 3])),
                     pack_bf16(__uint_as_float(o[8 * j + 4]), __uint_as_float(o[8 * j + 5])),
                     pack_bf16(__uint_as_float(o[8 * j + 6]), __uint_as_float(o[8 * j + 7])));
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 u = lds128_u(lds_base + i * 512);
            if (qb + rr0 + 8 * i < p.S) *reinterpret_cast<uint4*>(gp + i * gstep + half * 64) = u;
          }
          __syncwarp();
        }
      }
      tc_fence_before();
    };
    int n = 0;
    for (int it = it0; it < nitems; it += gs, ++n) {
      const int stage = n & 1;
      const int b = it / hp, h = (it - b * hp) * 2;
      VQ_TRR(n, 0);
      mbar_wait(&full_bar[stage], (n >> 1) & 1);     // header visible
      VQ_TRR(n, 1);
      mbar_wait(sfull_bar, n & 1);
      VQ_TRR(n, 2);
      tc_fence_after();
      float s[16], dp[16];
      {
        uint32_t t0[16], t1[16];
        const uint32_t ta = tmem_base + hsel * 64 + cq * 16 + lane_addr;
        tmem_ld_32x16(ta, t0);
        tmem_ld_32x16(ta + 128, t1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) { s[j] = __uint_as_float(t0[j]); dp[j] = __uint_as_float(t1[j]); }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sread_bar);
      const float* hs = hdr + stage * TCB_HDR_FLOATS;
      const float lse = hs[64 + 256 + r];
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 km = *reinterpret_cast<const float4*>(hs + cq * 16 + 4 * j4);
        s[4 * j4] += km.x; s[4 * j4 + 1] += km.y; s[4 * j4 + 2] += km.z; s[4 * j4 + 3] += km.w;
      }
      if (q < p.Lt && cq * 16 < p.Lt) {
        const float* bp = hs + 64 + hsel * 128 + (AT_S_TC - 1) - q + cq * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (cq * 16 + j < p.Lt) s[j] += bp[j];
      }
      uint32_t keep = 0xFFFFu;                     // bit j: probability (q, 16 cq + j) survived dropout
      if (p.drop_thr) {
        keep = 0u;
        const uint32_t pi0 = attn_pair_idx((uint32_t)(b * p.H + h + hsel), q, cq * 16);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t hsh = vq_hash_pair(p.seed, pi0 + j);
          keep |= ((hsh & 0xFFFFu) >= p.drop_thr ? 1u : 0u) << (2 * j);
          keep |= ((hsh >> 16) >= p.drop_thr ? 1u : 0u) << (2 * j + 1);
        }
      }
      const float ik = p.drop_thr ? p.drop_inv_keep : 1.f;
      float dpart = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        s[j] = __expf(s[j] - lse);                 // exp(-inf) = 0: masked keys, rows beyond S (lse = +inf)
        dp[j] = ((keep >> j) & 1u) ? dp[j] * ik : 0.f;
        dpart += s[j] * dp[j];
      }
      dxr[(n & 1) * 512 + cq * 128] = dpart;
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      VQ_TRR(n, 3);
      const float* dxp = dxr + (n & 1) * 512;
      const float dsum = (dxp[0] + dxp[128]) + (dxp[256] + dxp[384]);
      uint32_t pw[8], dw[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        const int j = 2 * x;
        pw[x] = pack_bf16(((keep >> j) & 1u) ? s[j] * ik : 0.f, ((keep >> (j + 1)) & 1u) ? s[j + 1] * ik : 0.f);
        dw[x] = pack_bf16(s[j] * (dp[j] - dsum), s[j + 1] * (dp[j + 1] - dsum));
      }
      // the previous item's output rows: its MMAs finished long ago; reading them here frees the dQ / dK / dV accumulators and
      // guarantees the P / dS tiles have been consumed before they are overwritten below
      if (n > 0) {
        epilogue(n - 1);
        if (p.d_rel_table) {
          if (cq == 3) {
            diagonals();
            asm volatile("bar.arrive 6, %0;" ::"n"(TCB_ROW_WARPS * 32) : "memory");
          } else {
            asm volatile("bar.sync 6, %0;" ::"n"(TCB_ROW_WARPS * 32) : "memory");     // the dS tile may be rewritten
          }
        }
      }
      VQ_TRR(n, 5);
      b_prev = b; h_prev = h;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t off = (uint32_t)((cq * 2 + c) ^ sw) << 4;
        sts128_u(sP_u + row_off + off, pw[4 * c], pw[4 * c + 1], pw[4 * c + 2], pw[4 * c + 3]);
        sts128_u(sdS_u + row_off + off, dw[4 * c], dw[4 * c + 1], dw[4 * c + 2], dw[4 * c + 3]);
      }
      (void)zero_off;                              // the off-diagonal halves were zeroed once and are never written
      fence_proxy_async();
      VQ_TRR(n, 6);
      __syncwarp();
      if (lane == 0) mbar_arrive(&pfull_bar[stage]);
      VQ_TRR(n, 7);
      if (p.d_rel_table && cq < 3) asm volatile("bar.arrive 5, %0;" ::"n"(TCB_ROW_WARPS * 32) : "memory");   // dS tile of item n: this warp's part is written
    }
    if (n > 0) epilogue(n - 1);
    if (p.d_rel_table) {
      if (n > 0 && cq == 3) diagonals();
      asm volatile("bar.sync 7, %0;" ::"n"(TCB_ROW_WARPS * 32) : "memory");
      for (int i = threadIdx.x - TCB_FIRST_ROW_WARP * 32; i < TCB_DTAB_FLOATS; i += TCB_ROW_WARPS * 32) {
        const float v = dtab[i];
        const int bucket = i >> 4, head = i & 15;
        if (v != 0.f && head < p.H) atomicAdd(&p.d_rel_table[bucket * p.H + head], v);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TCB_ISSUER_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

// VQACL_ATTN_TC=0 falls back to the mma.sync kernel of attention.cu (A/B measurements)
static int g_attn_tc = [] { const char* e = getenv("VQACL_ATTN_TC"); return (e && e[0] == '0') ? 0 : 1; }();

// true when the tcgen05 kernel takes this problem (encoder form: square, 33..64 positions, text-corner bias, even head count)
bool attn_tc_eligible(const AttnArgs& a) {
  return g_attn_tc && a.rel_mode == 1 && a.Sq == a.Sk && a.Sq > 32 && a.Sq <= 64 && a.Lt <= 32 && (a.H & 1) == 0 && !a.causal && !a.q_off &&
         !a.q_bstride && !a.k_bstride && !a.v_bstride && !a.o_bstride && a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(a.q) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.k) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(a.v) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.o) & 15) == 0 && a.ldo % 8 == 0;
}

int attn_enc_fwd_tc(const AttnArgs& a, const AttnBuckets& bk, cudaStream_t stream) {
  CUtensorMap tq, tk, tv;
  const uint64_t rows = (uint64_t)a.B * a.Sq, cols = (uint64_t)a.H * 64;
  if (make_tmap_bf16_2d(&tq, a.q, cols, rows, a.ldq, 64, 64)) return 1;
  if (make_tmap_bf16_2d(&tk, a.k, cols, rows, a.ldk, 64, 64)) return 1;
  if (make_tmap_bf16_2d(&tv, a.v, cols, rows, a.ldv, 64, 64)) return 1;
  static bool attr = false;
  if (!attr) {
    VQ_CUDA(cudaFuncSetAttribute(attn_enc_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TCF_SMEM_BYTES));
    attr = true;
  }
  AttnTcArgs p{};
  p.o = a.o; p.ldo = a.ldo; p.lse = a.lse; p.B = a.B; p.H = a.H; p.S = a.Sq; p.Lt = a.Lt; p.rel_table = a.rel_table; p.keymask = a.keymask;
  p.drop_thr = a.drop_thr; p.drop_inv_keep = a.drop_inv_keep; p.seed = a.seed;
  const int items = a.B * (a.H / 2);
  const int cap = TCF_CTAS_PER_SM * num_sms();
  const int grid = items < cap ? items : cap;
  (void)vq_launch(attn_enc_fwd_tc_kernel, dim3(grid), dim3(TCF_THREADS), (size_t)TCF_SMEM_BYTES, stream, tq, tk, tv, p, bk);
  VQ_LAUNCH_CHECK();
  return 0;
}

bool attn_tc_bwd_eligible(const AttnArgs& a) {
  static const int on = [] { const char* e = getenv("VQACL_ATTN_TC_BWD"); return (e && e[0] == '0') ? 0 : 1; }();   // =0: mma.sync backward (A/B)
  return on && a.rel_mode == 1 && a.Sq == a.Sk && a.Sq > 32 && a.Sq <= 64 && a.Lt <= 32 && (a.H & 1) == 0 && !a.causal && a.ldq % 8 == 0 &&
         a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(a.q) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(a.k) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.v) & 15) == 0 && a.H <= 16 && a.lddq % 8 == 0 && a.lddk % 8 == 0 && a.lddv % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(a.dO) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.dq) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(a.dk) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.dv) & 15) == 0;
}

int attn_enc_bwd_tc(const AttnArgs& a, const AttnBuckets& bk, cudaStream_t stream) {
  CUtensorMap tq, tk, tv, tdo;
  const uint64_t rows = (uint64_t)a.B * a.Sq, cols = (uint64_t)a.H * 64;
  if (make_tmap_bf16_2d(&tq, a.q, cols, rows, a.ldq, 64, 64)) return 1;
  if (make_tmap_bf16_2d(&tk, a.k, cols, rows, a.ldk, 64, 64)) return 1;
  if (make_tmap_bf16_2d(&tv, a.v, cols, rows, a.ldv, 64, 64)) return 1;
  if (make_tmap_bf16_2d(&tdo, a.dO, cols, rows, a.ldo, 64, 64)) return 1;
  static bool attr = false;
  if (!attr) {
    VQ_CUDA(cudaFuncSetAttribute(attn_enc_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TCB_SMEM_BYTES));
    attr = true;
  }
  AttnTcBwdArgs p{};
  p.dq = a.dq; p.dk = a.dk; p.dv = a.dv; p.lddq = a.lddq; p.lddk = a.lddk; p.lddv = a.lddv; p.lse = a.lse;
  p.B = a.B; p.H = a.H; p.S = a.Sq; p.Lt = a.Lt; p.rel_table = a.rel_table; p.keymask = a.keymask; p.d_rel_table = a.d_rel_table;
  p.drop_thr = a.drop_thr; p.drop_inv_keep = a.drop_inv_keep; p.seed = a.seed;
  const int items = a.B * (a.H / 2);
  const int grid = items < num_sms() ? items : num_sms();
  (void)vq_launch(attn_enc_bwd_tc_kernel, dim3(grid), dim3(TCB_THREADS), (size_t)TCB_SMEM_BYTES, stream, tq, tk, tv, tdo, p, bk);
  VQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace vq

#ifdef VQ_ATTN_TRACE
// debug builds only (not declared in include/vqacl_b200.h): installs the device buffer [4][16][8] of int64 clock stamps
extern "C" int vqacl_debug_attn_trace(void* buf) {
  VQ_CUDA(cudaMemcpyToSymbol(vq::g_attn_trace, &buf, sizeof(buf)));
  return 0;
}
#endif
