// Fused T5 attention for the VL-T5 shapes (S <= 64 keys/queries, d_kv = 64): one CTA per (batch, head) holds the whole
// Q/K/V problem in shared memory. Scores are NOT scaled by 1/sqrt(d) (T5), the relative-position bias is read from the
// [num_buckets, H] embedding table through a host-precomputed rel->bucket map, and the additive masks of HF 4.2.1
// (-10000 key padding / causal, -1e9 cross) are applied in-kernel, so the [B,H,S,S] bias tensor the reference
// materialises (modeling_t5_our.py:258-273) never exists. Backward recomputes P from the saved log-sum-exp.
//
// Reference math: HF T5Attention (hf5.5 modeling_t5.py:277-338 == 4.2.1): softmax(QK^T + bias) V with dropout on P.
// The QK^T / PV cores are ~1 % of the step's FLOPs: they run on mma.sync m16n8k16 bf16 fragments;
// the tcgen05 path is reserved for the projections that carry the FLOPs.
#include "common.cuh"
#include "ops.h"

namespace vq {

constexpr int AT_S = 64;      // max queries / keys per problem
constexpr int AT_D = 64;      // head dim
constexpr int AT_P = 72;      // smem pitch (elements): 144 B rows -> conflict-free ldmatrix
constexpr int AT_THREADS = 128;

VQ_DEVINL void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
VQ_DEVINL void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
VQ_DEVINL void ldsm_x2(uint32_t (&r)[2], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(smem_u32(p)));
}
VQ_DEVINL void ldsm_x2_t(uint32_t (&r)[2], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(smem_u32(p)));
}
VQ_DEVINL void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// load a [rows<=64, 64] bf16 head slice into smem, 16-byte vectors; rows [rows, round_up(rows,16)) are zero-filled and
// rows beyond that are left untouched (every consumer loop is bounded by the same round-up)
VQ_DEVINL void load_head(__nv_bfloat16 (*dst)[AT_P], const __nv_bfloat16* src, int ld, int rows) {
  const int rows16 = (rows + 15) & ~15;
  for (int idx = threadIdx.x; idx < rows16 * (AT_D / 8); idx += AT_THREADS) {
    const int r = idx >> 3, c = (idx & 7) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < rows) v = *reinterpret_cast<const uint4*>(src + (size_t)r * ld + c);
    *reinterpret_cast<uint4*>(&dst[r][c]) = v;
  }
}

// A fragment (m16 x k16) of row-major X[m][k] at (m0, k0)
VQ_DEVINL void frag_a(uint32_t (&a)[4], const __nv_bfloat16 (*X)[AT_P], int m0, int k0, int lane) {
  const int mi = lane >> 3, r = lane & 7;
  ldsm_x4(a, &X[m0 + (mi & 1) * 8 + r][k0 + (mi >> 1) * 8]);
}
// A fragment of X^T where X is stored [k][m] row-major
VQ_DEVINL void frag_a_t(uint32_t (&a)[4], const __nv_bfloat16 (*X)[AT_P], int m0, int k0, int lane) {
  const int mi = lane >> 3, r = lane & 7;
  ldsm_x4_t(a, &X[k0 + (mi >> 1) * 8 + r][m0 + (mi & 1) * 8]);
}
// B fragment (k16 x n8) where B[k][n] is stored as X[n][k] row-major
VQ_DEVINL void frag_b(uint32_t (&b)[2], const __nv_bfloat16 (*X)[AT_P], int n0, int k0, int lane) {
  const int l = lane & 15, mi = l >> 3, r = l & 7;
  ldsm_x2(b, &X[n0 + r][k0 + mi * 8]);
}
// B fragment where B[k][n] is stored as X[k][n] row-major
VQ_DEVINL void frag_b_t(uint32_t (&b)[2], const __nv_bfloat16 (*X)[AT_P], int n0, int k0, int lane) {
  const int l = lane & 15, mi = l >> 3, r = l & 7;
  ldsm_x2_t(b, &X[k0 + mi * 8 + r][n0]);
}

// dropout pair index of probabilities (q, k), (q, k+1) of problem `blk` (k even)
VQ_DEVINL uint32_t attn_pair_idx(uint32_t blk, int q, int k) { return ((blk * AT_S + (uint32_t)q) * AT_S + (uint32_t)k) >> 1; }

struct AttnSmemFwd {
  __nv_bfloat16 q[AT_S][AT_P], k[AT_S][AT_P], v[AT_S][AT_P];
  float bias[2 * AT_S];
  float kmask[AT_S];
};

// scores for the warp's 16 query rows vs the keys, bias/mask applied; keys >= Sk get -inf.
// Only the ceil(Sk/8) key tiles that exist are multiplied (decoder self-attention has Sk = T <= 10).
VQ_DEVINL void scores_tile(float (&s)[8][4], const __nv_bfloat16 (*sq)[AT_P], const __nv_bfloat16 (*sk)[AT_P],
                           const float* sbias, const float* skmask, const AttnArgs& p, int m0, int lane) {
  const int nkt = (p.Sk + 7) >> 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) s[nt][i] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    frag_a(a, sq, m0, kk * 16, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt < nkt) {
        uint32_t b[2];
        frag_b(b, sk, nt * 8, kk * 16, lane);
        mma16816(s[nt], a, b);
      }
    }
  }
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = m0 + g + (i >> 1) * 8 + p.q_off;   // absolute query position
      const int kj = nt * 8 + 2 * t + (i & 1);
      float x = s[nt][i];
      if (kj >= p.Sk) {
        x = -INFINITY;
      } else {
        if (p.rel_mode == 2 || (p.rel_mode == 1 && qi < p.Lt && kj < p.Lt)) x += sbias[min(max(kj - qi, -(AT_S - 1)), AT_S - 1) + (AT_S - 1)];
        x += skmask[kj];
        if (p.causal && kj > qi) x += -10000.0f;
      }
      s[nt][i] = x;
    }
}

VQ_DEVINL void load_bias_mask(float* sbias, float* skmask, const AttnArgs& p, int b, int h) {
  if (p.rel_mode)
    for (int r = threadIdx.x; r < 2 * AT_S - 1; r += AT_THREADS) sbias[r] = p.rel_table[p.rel_bucket[r] * p.H + h];
  for (int j = threadIdx.x; j < AT_S; j += AT_THREADS) skmask[j] = (p.keymask && j < p.Sk) ? p.keymask[(size_t)b * p.Sk + j] : 0.f;
}

__global__ void __launch_bounds__(AT_THREADS, 4) attn_fwd_kernel(const AttnArgs p) {
  extern __shared__ uint8_t at_smem_raw[];
  AttnSmemFwd& sm = *reinterpret_cast<AttnSmemFwd*>(at_smem_raw);
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  load_head(sm.q, p.q + (size_t)b * (p.q_bstride ? p.q_bstride : (long long)p.Sq * p.ldq) + h * AT_D, p.ldq, p.Sq);
  load_head(sm.k, p.k + (size_t)b * (p.k_bstride ? p.k_bstride : (long long)p.Sk * p.ldk) + h * AT_D, p.ldk, p.Sk);
  load_head(sm.v, p.v + (size_t)b * (p.v_bstride ? p.v_bstride : (long long)p.Sk * p.ldv) + h * AT_D, p.ldv, p.Sk);
  load_bias_mask(sm.bias, sm.kmask, p, b, h);
  __syncthreads();
  const int m0 = warp * 16;
  if (m0 >= p.Sq) return;
  float s[8][4];
  scores_tile(s, sm.q, sm.k, sm.bias, sm.kmask, p, m0, lane);
  const int g = lane >> 2, t = lane & 3;
  // row-wise softmax (rows g and g+8 of this warp's tile), fp32
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) mx[i >> 1] = fmaxf(mx[i >> 1], s[nt][i]);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
  }
  float sum[2] = {0.f, 0.f};
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float e = __expf(s[nt][i] - mx[i >> 1]);
      s[nt][i] = e;
      sum[i >> 1] += e;
    }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
    sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
  }
  const float inv[2] = {1.f / sum[0], 1.f / sum[1]};
  if (t == 0 && p.lse) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int qi = m0 + g + r * 8;
      if (qi < p.Sq) p.lse[((size_t)b * p.H + h) * p.Sq + qi] = mx[r] + __logf(sum[r]);
    }
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float s0 = inv[r], s1 = inv[r];
      if (p.drop_thr) {
        float d0, d1;
        vq_dropout_pair(p.seed, attn_pair_idx(blockIdx.x, m0 + g + r * 8, nt * 8 + 2 * t), p.drop_thr, p.drop_inv_keep, d0, d1);
        s0 *= d0; s1 *= d1;
      }
      s[nt][2 * r] *= s0;
      s[nt][2 * r + 1] *= s1;
    }
  // O = P V over the ceil(Sk/16) key blocks that exist
  const int nkk = (p.Sk + 15) >> 4;
  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[nt][i] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    if (kk < nkk) {
      uint32_t a[4];
      a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        uint32_t bb[2];
        frag_b_t(bb, sm.v, nt * 8, kk * 16, lane);
        mma16816(o[nt], a, bb);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qi = m0 + g + r * 8;
    if (qi < p.Sq) {
      __nv_bfloat16* dst = p.o + (size_t)b * (p.o_bstride ? p.o_bstride : (long long)p.Sq * p.ldo) + (size_t)qi * p.ldo + h * AT_D + 2 * t;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<uint32_t*>(dst + nt * 8) = pack_bf16(o[nt][2 * r], o[nt][2 * r + 1]);
    }
  }
}

struct AttnSmemBwd {
  __nv_bfloat16 q[AT_S][AT_P], k[AT_S][AT_P], v[AT_S][AT_P], dO[AT_S][AT_P], P[AT_S][AT_P], dS[AT_S][AT_P];
  float bias[2 * AT_S];
  float kmask[AT_S];
  float dbucket[64];
};

__global__ void __launch_bounds__(AT_THREADS, 3) attn_bwd_kernel(const AttnArgs p) {
  extern __shared__ uint8_t at_smem_raw[];
  AttnSmemBwd& sm = *reinterpret_cast<AttnSmemBwd*>(at_smem_raw);
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  load_head(sm.q, p.q + (size_t)b * p.Sq * p.ldq + h * AT_D, p.ldq, p.Sq);
  load_head(sm.k, p.k + (size_t)b * p.Sk * p.ldk + h * AT_D, p.ldk, p.Sk);
  load_head(sm.v, p.v + (size_t)b * p.Sk * p.ldv + h * AT_D, p.ldv, p.Sk);
  load_head(sm.dO, p.dO + (size_t)b * p.Sq * p.ldo + h * AT_D, p.ldo, p.Sq);
  load_bias_mask(sm.bias, sm.kmask, p, b, h);
  if (threadIdx.x < 64) sm.dbucket[threadIdx.x] = 0.f;
  __syncthreads();

  const int nkt = (p.Sk + 7) >> 3;    // key tiles of 8 that exist
  const int nkk = (p.Sk + 15) >> 4;   // key blocks of 16
  const int nqk = (p.Sq + 15) >> 4;   // query blocks of 16
  // ---- phase 1: this warp owns 16 query rows (warps past the last query block only take part in phase 2) ----
  const int m0 = warp * 16;
  if (m0 < p.Sq) {
    float s[8][4];
    scores_tile(s, sm.q, sm.k, sm.bias, sm.kmask, p, m0, lane);
    float lse[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int qi = m0 + g + r * 8;
      lse[r] = qi < p.Sq ? p.lse[((size_t)b * p.H + h) * p.Sq + qi] : INFINITY;  // rows >= Sq -> P = 0
    }
    // dPd = dO V^T
    float dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) dp[nt][i] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      frag_a(a, sm.dO, m0, kk * 16, lane);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt < nkt) {
          uint32_t bb[2];
          frag_b(bb, sm.v, nt * 8, kk * 16, lane);
          mma16816(dp[nt], a, bb);
        }
      }
    }
    float dsum[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int qi = m0 + g + r * 8, kj = nt * 8 + 2 * t;
        float sc0 = 1.f, sc1 = 1.f;
        if (p.drop_thr) vq_dropout_pair(p.seed, attn_pair_idx(blockIdx.x, qi, kj), p.drop_thr, p.drop_inv_keep, sc0, sc1);
        const float p0 = __expf(s[nt][2 * r] - lse[r]), p1 = __expf(s[nt][2 * r + 1] - lse[r]);  // exp(-inf) = 0: masked keys / padded rows
        const float d0 = dp[nt][2 * r] * sc0, d1 = dp[nt][2 * r + 1] * sc1;                          // dP
        s[nt][2 * r] = p0; s[nt][2 * r + 1] = p1;
        dp[nt][2 * r] = d0; dp[nt][2 * r + 1] = d1;
        dsum[r] += p0 * d0 + p1 * d1;
        *reinterpret_cast<uint32_t*>(&sm.P[qi][kj]) = pack_bf16(p0 * sc0, p1 * sc1);                  // dropped P, for dV
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      dsum[r] += __shfl_xor_sync(0xffffffffu, dsum[r], 1);
      dsum[r] += __shfl_xor_sync(0xffffffffu, dsum[r], 2);
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float ds0 = s[nt][2 * r] * (dp[nt][2 * r] - dsum[r]), ds1 = s[nt][2 * r + 1] * (dp[nt][2 * r + 1] - dsum[r]);
        s[nt][2 * r] = ds0; s[nt][2 * r + 1] = ds1;
        *reinterpret_cast<uint32_t*>(&sm.dS[m0 + g + r * 8][nt * 8 + 2 * t]) = pack_bf16(ds0, ds1);
      }
    // dQ = dS K   (B[k=key][n=d] stored [key][d] -> transposed fragment loads)
    float dq[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) dq[nt][i] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk < nkk) {
        uint32_t a[4];
        a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
        a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
        a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          uint32_t bb[2];
          frag_b_t(bb, sm.k, nt * 8, kk * 16, lane);
          mma16816(dq[nt], a, bb);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int qi = m0 + g + r * 8;
      if (qi < p.Sq) {
        __nv_bfloat16* dst = p.dq + ((size_t)b * p.Sq + qi) * p.lddq + h * AT_D + 2 * t;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<uint32_t*>(dst + nt * 8) = pack_bf16(dq[nt][2 * r], dq[nt][2 * r + 1]);
      }
    }
  }
  __syncthreads();
  // ---- relative-position-bias gradient: d table[bucket(k - q)] += dS[q][k]; one thread per diagonal of the biased
  //      region sums its <= 64 entries from smem, then one smem atomic per diagonal and one global atomic per bucket ----
  if (p.d_rel_table) {
    const int nq = p.rel_mode == 1 ? min(p.Sq, p.Lt) : p.Sq;
    const int nk = p.rel_mode == 1 ? min(p.Sk, p.Lt) : p.Sk;
    const int ndiag = nq + nk - 1;
    for (int dgi = threadIdx.x; dgi < ndiag; dgi += AT_THREADS) {
      const int rel = dgi - (nq - 1);   // k - q
      float acc = 0.f;
      for (int qi = max(0, -rel); qi < nq && qi + rel < nk; ++qi) acc += __bfloat162float(sm.dS[qi][qi + rel]);
      atomicAdd(&sm.dbucket[p.rel_bucket[rel + (AT_S - 1)]], acc);
    }
  }
  // ---- phase 2: this warp owns 16 key rows: dV = Pd^T dO, dK = dS^T Q (contraction over the query blocks that exist) ----
  if (m0 < p.Sk) {
    float dv[8][4], dk[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) dv[nt][i] = dk[nt][i] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk < nqk) {
        uint32_t ap[4], as[4];
        frag_a_t(ap, sm.P, m0, kk * 16, lane);
        frag_a_t(as, sm.dS, m0, kk * 16, lane);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          uint32_t b1[2], b2[2];
          frag_b_t(b1, sm.dO, nt * 8, kk * 16, lane);
          frag_b_t(b2, sm.q, nt * 8, kk * 16, lane);
          mma16816(dv[nt], ap, b1);
          mma16816(dk[nt], as, b2);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int kj = m0 + g + r * 8;
      if (kj < p.Sk) {
        __nv_bfloat16* dstk = p.dk + ((size_t)b * p.Sk + kj) * p.lddk + h * AT_D + 2 * t;
        __nv_bfloat16* dstv = p.dv + ((size_t)b * p.Sk + kj) * p.lddv + h * AT_D + 2 * t;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          *reinterpret_cast<uint32_t*>(dstk + nt * 8) = pack_bf16(dk[nt][2 * r], dk[nt][2 * r + 1]);
          *reinterpret_cast<uint32_t*>(dstv + nt * 8) = pack_bf16(dv[nt][2 * r], dv[nt][2 * r + 1]);
        }
      }
    }
  }
  if (p.d_rel_table) {
    __syncthreads();
    if (threadIdx.x < 64) {
      const float v = sm.dbucket[threadIdx.x];
      if (v != 0.f) atomicAdd(&p.d_rel_table[threadIdx.x * p.H + h], v);
    }
  }
}

static int check_args(const AttnArgs& a) {
  VQ_CHECK(a.Sq >= 1 && a.Sq <= AT_S && a.Sk >= 1 && a.Sk <= AT_S, "attention: Sq=%d Sk=%d must be in [1,%d]", a.Sq, a.Sk, AT_S);
  VQ_CHECK(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0, "attention: pitches must be multiples of 8");
  VQ_CHECK(a.rel_mode == 0 || (a.rel_table && a.rel_bucket), "attention: rel_mode needs rel_table and rel_bucket");
  return 0;
}

int attn_fwd(const AttnArgs& a, cudaStream_t stream) {
  if (check_args(a)) return 1;
  static bool attr = false;
  if (!attr) {
    VQ_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AttnSmemFwd)));
    attr = true;
  }
  attn_fwd_kernel<<<a.B * a.H, AT_THREADS, sizeof(AttnSmemFwd), stream>>>(a);
  VQ_LAUNCH_CHECK();
  return 0;
}

int attn_bwd(const AttnArgs& a, cudaStream_t stream) {
  if (check_args(a)) return 1;
  VQ_CHECK(a.lse && a.dO && a.dq && a.dk && a.dv, "attention bwd: missing pointers");
  VQ_CHECK(!a.q_bstride && !a.k_bstride && !a.v_bstride && !a.o_bstride && !a.q_off, "attention bwd: strided / offset form is forward-only");
  VQ_CHECK(a.lddq % 8 == 0 && a.lddk % 8 == 0 && a.lddv % 8 == 0, "attention bwd: pitches must be multiples of 8");
  static bool attr = false;
  if (!attr) {
    VQ_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AttnSmemBwd)));
    attr = true;
  }
  attn_bwd_kernel<<<a.B * a.H, AT_THREADS, sizeof(AttnSmemBwd), stream>>>(a);
  VQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace vq
