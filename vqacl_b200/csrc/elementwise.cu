// HBM-bound row kernels of the VL-T5 hot path: T5 RMSNorm fwd/bwd (hf5.5 modeling_t5.py:46-68), token-embedding
// gather / scatter-add, the VisualEmbedding tail (modeling_t5_our.py:93-143), mask construction and casts.
// One warp owns one 768-wide row: lane l holds the float4 chunks {l, l+32, ..., l+160} -> fully coalesced 512 B
// wavefronts, warp-shuffle reductions, no shared memory on the forward paths.
#include "ops.h"

#include <stdlib.h>

namespace vq {

constexpr int RW_CHUNKS = DM / 4 / 32;  // 6 float4 chunks per lane
constexpr int ROW_WARPS = 8;            // warps per CTA for the row kernels

VQ_DEVINL int out_row_of(int r, int in_rpb, int out_rpb) { return in_rpb > 0 ? (r / in_rpb) * out_rpb + r % in_rpb : r; }

VQ_DEVINL void load_row_f32(float (&v)[RW_CHUNKS][4], const float* row, int lane) {
#pragma unroll
  for (int j = 0; j < RW_CHUNKS; ++j) {
    const float4 t = *reinterpret_cast<const float4*>(row + (lane + 32 * j) * 4);
    v[j][0] = t.x; v[j][1] = t.y; v[j][2] = t.z; v[j][3] = t.w;
  }
}
VQ_DEVINL void load_row_bf16(float (&v)[RW_CHUNKS][4], const __nv_bfloat16* row, int lane) {
#pragma unroll
  for (int j = 0; j < RW_CHUNKS; ++j) {
    const uint2 t = *reinterpret_cast<const uint2*>(row + (lane + 32 * j) * 4);
    const float2 a = unpack_bf16(t.x), b = unpack_bf16(t.y);
    v[j][0] = a.x; v[j][1] = a.y; v[j][2] = b.x; v[j][3] = b.y;
  }
}
VQ_DEVINL void store_row_f32(float* row, const float (&v)[RW_CHUNKS][4], int lane) {
#pragma unroll
  for (int j = 0; j < RW_CHUNKS; ++j)
    *reinterpret_cast<float4*>(row + (lane + 32 * j) * 4) = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
}
VQ_DEVINL void store_row_bf16(__nv_bfloat16* row, const float (&v)[RW_CHUNKS][4], int lane) {
#pragma unroll
  for (int j = 0; j < RW_CHUNKS; ++j)
    *reinterpret_cast<uint2*>(row + (lane + 32 * j) * 4) = make_uint2(pack_bf16(v[j][0], v[j][1]), pack_bf16(v[j][2], v[j][3]));
}
VQ_DEVINL float row_sumsq(const float (&v)[RW_CHUNKS][4]) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < RW_CHUNKS; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) s += v[j][i] * v[j][i];
  return warp_sum(s);
}
VQ_DEVINL void apply_dropout(float (&v)[RW_CHUNKS][4], const Dropout& d, uint64_t row_elem0, int lane) {
  if (!d.thr) return;
  const uint32_t pair0 = (uint32_t)(row_elem0 >> 1);   // row_elem0 is a multiple of 4
#pragma unroll
  for (int j = 0; j < RW_CHUNKS; ++j) {
    const uint32_t pi = pair0 + (uint32_t)(lane + 32 * j) * 2u;
    float s0, s1, s2, s3;
    vq_dropout_pair(d.seed, pi, d.thr, d.inv_keep, s0, s1);
    vq_dropout_pair(d.seed, pi + 1, d.thr, d.inv_keep, s2, s3);
    v[j][0] *= s0; v[j][1] *= s1; v[j][2] *= s2; v[j][3] *= s3;
  }
}

// ------------------------------------------------------------------------------------------------ cast
__global__ void cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n8) {
  vq_pdl_trigger();
  vq_pdl_wait();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(src)[2 * i], b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    reinterpret_cast<uint4*>(dst)[i] = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
}
int cast_f32_to_bf16(const float* src, __nv_bfloat16* dst, size_t n, cudaStream_t stream) {
  VQ_CHECK(n % 8 == 0, "cast: n=%zu must be a multiple of 8", n);
  if (n == 0) return 0;
  const size_t n8 = n / 8;
  const int blocks = (int)((n8 + 255) / 256 < (size_t)num_sms() * 8 ? (n8 + 255) / 256 : (size_t)num_sms() * 8);
  (void)vq_launch(cast_kernel, dim3(blocks), dim3(256), 0, stream, src, dst, n8);
  VQ_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------ RMSNorm fwd
__global__ void __launch_bounds__(ROW_WARPS * 32) rmsnorm_fwd_kernel(const RmsFwdArgs a) {
  vq_pdl_trigger();
  vq_pdl_wait();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (r >= a.M) return;
  float v[RW_CHUNKS][4], w[RW_CHUNKS][4];
  if (a.parts) {
    float t[RW_CHUNKS][4];
    load_row_f32(v, a.parts + (size_t)r * DM, lane);
    for (int s = 1; s < a.n_parts; ++s) {
      load_row_f32(t, a.parts + (size_t)s * a.part_stride + (size_t)r * DM, lane);
#pragma unroll
      for (int j = 0; j < RW_CHUNKS; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) v[j][i] += t[j][i];
    }
    apply_dropout(v, a.resid_drop, (uint64_t)r * DM, lane);
    load_row_f32(t, a.resid + (size_t)r * DM, lane);
#pragma unroll
    for (int j = 0; j < RW_CHUNKS; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) v[j][i] += t[j][i];
    store_row_f32(a.x_out + (size_t)r * DM, v, lane);
  } else {
    load_row_f32(v, a.x + (size_t)r * DM, lane);
  }
  load_row_f32(w, a.w, lane);
  const float rstd = rsqrtf(row_sumsq(v) / DM + a.eps) * a.scale;
#pragma unroll
  for (int j = 0; j < RW_CHUNKS; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) v[j][i] = v[j][i] * rstd * w[j][i];
  apply_dropout(v, a.drop, (uint64_t)r * DM, lane);
  const int ro = out_row_of(r, a.in_rpb, a.out_rpb);
  if (a.y_bf16) store_row_bf16(a.y_bf16 + (size_t)ro * a.ld_bf16, v, lane);
  if (a.y_f32) store_row_f32(a.y_f32 + (size_t)r * a.ld_f32, v, lane);
}
int rmsnorm_fwd(const RmsFwdArgs& a, cudaStream_t stream) {
  if (a.M <= 0) return 0;
  (void)vq_launch(rmsnorm_fwd_kernel, dim3((a.M + ROW_WARPS - 1) / ROW_WARPS), dim3(ROW_WARPS * 32), 0, stream, a);
  VQ_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------ RMSNorm bwd
__global__ void __launch_bounds__(ROW_WARPS * 32) rmsnorm_bwd_kernel(const RmsBwdArgs a) {
  vq_pdl_trigger();
  vq_pdl_wait();
  __shared__ float s_dw[ROW_WARPS][DM];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float w[RW_CHUNKS][4], dwacc[RW_CHUNKS][4];
  load_row_f32(w, a.w, lane);
#pragma unroll
  for (int j = 0; j < RW_CHUNKS; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) dwacc[j][i] = 0.f;
  for (int r = blockIdx.x * ROW_WARPS + warp; r < a.M; r += gridDim.x * ROW_WARPS) {
    float x[RW_CHUNKS][4], dy[RW_CHUNKS][4];
    load_row_f32(x, a.x + (size_t)r * DM, lane);
    const int rd = out_row_of(r, a.in_rpb, a.out_rpb);
    if (a.dn_f32) {
      load_row_f32(dy, a.dn_f32 + (size_t)rd * a.ld_dn, lane);
      if (a.dn_zero) {
        float* zp = const_cast<float*>(a.dn_f32) + (size_t)rd * a.ld_dn;
#pragma unroll
        for (int j = 0; j < RW_CHUNKS; ++j) *reinterpret_cast<float4*>(zp + (lane + 32 * j) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      load_row_bf16(dy, a.dn + (size_t)rd * a.ld_dn, lane);
    }
    if (a.dn2) {
      float d2[RW_CHUNKS][4];
      load_row_bf16(d2, a.dn2 + (size_t)rd * a.ld_dn2, lane);
#pragma unroll
      for (int j = 0; j < RW_CHUNKS; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) dy[j][i] += d2[j][i];
    }
    if (a.bc_q) {
      const int bb = r / a.bc_S, tt = r - bb * a.bc_S;
      const bool qs = tt < a.bc_split;
      const float cf = qs ? a.bc_g[0] * a.bc_cq : a.bc_g[1] * a.bc_cv;
      float bc[RW_CHUNKS][4];
      load_row_f32(bc, (qs ? a.bc_q : a.bc_v) + (size_t)bb * DM, lane);
#pragma unroll
      for (int j = 0; j < RW_CHUNKS; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) dy[j][i] += cf * bc[j][i];
    }
    apply_dropout(dy, a.own, (uint64_t)r * DM, lane);
    const float rstd = rsqrtf(row_sumsq(x) / DM + a.eps);
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < RW_CHUNKS; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        x[j][i] *= rstd;                      // xhat
        dy[j][i] *= a.scale;                  // d(y / scale)
        dwacc[j][i] += dy[j][i] * x[j][i];
        dy[j][i] *= w[j][i];                  // dxhat
        dot += dy[j][i] * x[j][i];
      }
    dot = warp_sum(dot) / DM;
    float gout[RW_CHUNKS][4];
    if (a.g_in) load_row_f32(gout, a.g_in + (size_t)r * DM, lane);
#pragma unroll
    for (int j = 0; j < RW_CHUNKS; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dx = rstd * (dy[j][i] - x[j][i] * dot);
        gout[j][i] = a.g_in ? gout[j][i] + dx : dx;
      }
    if (a.g_out) store_row_f32(a.g_out + (size_t)r * DM, gout, lane);
    if (a.gb_out) {
      apply_dropout(gout, a.consumer, (uint64_t)r * a.consumer_cols, lane);
      store_row_bf16(a.gb_out + (size_t)r * DM, gout, lane);
    }
  }
  if (a.dw) {
#pragma unroll
    for (int j = 0; j < RW_CHUNKS; ++j)
      *reinterpret_cast<float4*>(&s_dw[warp][(lane + 32 * j) * 4]) = make_float4(dwacc[j][0], dwacc[j][1], dwacc[j][2], dwacc[j][3]);
    __syncthreads();
    for (int c = threadIdx.x; c < DM; c += ROW_WARPS * 32) {
      float s = 0.f;
#pragma unroll
      for (int wv = 0; wv < ROW_WARPS; ++wv) s += s_dw[wv][c];
      atomicAdd(&a.dw[c], s);
    }
  }
}
int rmsnorm_bwd(const RmsBwdArgs& a, cudaStream_t stream) {
  if (a.M <= 0) return 0;
  int blocks = (a.M + ROW_WARPS - 1) / ROW_WARPS;
  static const int per_sm = [] { const char* ev = getenv("VQACL_RMS_BWD_CTAS_PER_SM"); return ev && atoi(ev) > 0 ? atoi(ev) : 2; }();
  const int cap = num_sms() * per_sm;
  if (blocks > cap) blocks = cap;
  (void)vq_launch(rmsnorm_bwd_kernel, dim3(blocks), dim3(ROW_WARPS * 32), 0, stream, a);
  VQ_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------ embeddings
__global__ void __launch_bounds__(ROW_WARPS * 32)
embed_fwd_kernel(const int64_t* __restrict__ ids, int B, int L, const float* __restrict__ table, float* __restrict__ x, int S,
                 int row0, Dropout drop, int vocab, int* __restrict__ err) {
  vq_pdl_trigger();
  vq_pdl_wait();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (r >= B * L) return;
  const int b = r / L, i = r % L;
  float v[RW_CHUNKS][4];
  int64_t id = ids[r];
  if (id < 0 || id >= vocab) {   // torch raises IndexError here; we read row 0 and raise on the host at the next error check
    if (lane == 0 && err) atomicOr(err, 1);
    id = 0;
  }
  load_row_f32(v, table + (size_t)id * DM, lane);
  const size_t orow = (size_t)b * S + row0 + i;
  apply_dropout(v, drop, orow * DM, lane);
  store_row_f32(x + orow * DM, v, lane);
}
int embed_fwd(const int64_t* ids, int B, int L, const float* table, float* x, int S, int row0, Dropout drop, int vocab, int* err,
              cudaStream_t stream) {
  if (B * L <= 0) return 0;
  (void)vq_launch(embed_fwd_kernel, dim3((B * L + ROW_WARPS - 1) / ROW_WARPS), dim3(ROW_WARPS * 32), 0, stream, ids, B, L, table, x, S, row0, drop,
                  vocab, err);
  VQ_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(ROW_WARPS * 32)
embed_bwd_kernel(const int64_t* __restrict__ ids, int B, int L, const float* __restrict__ g, int S, int row0,
                 float* __restrict__ dtable, Dropout drop, int vocab) {
  vq_pdl_trigger();
  vq_pdl_wait();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (r >= B * L) return;
  if (ids[r] < 0 || ids[r] >= vocab) return;   // flagged by embed_fwd; never write outside the table's gradient
  const int b = r / L, i = r % L;
  const size_t grow = (size_t)b * S + row0 + i;
  float v[RW_CHUNKS][4];
  load_row_f32(v, g + grow * DM, lane);
  apply_dropout(v, drop, grow * DM, lane);
  float* dst = dtable + (size_t)ids[r] * DM;
#pragma unroll
  for (int j = 0; j < RW_CHUNKS; ++j)
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + (lane + 32 * j) * 4), "f"(v[j][0]), "f"(v[j][1]),
                 "f"(v[j][2]), "f"(v[j][3])
                 : "memory");
}
int embed_bwd(const int64_t* ids, int B, int L, const float* g, int S, int row0, float* dtable, Dropout drop, int vocab,
              cudaStream_t stream) {
  if (B * L <= 0) return 0;
  (void)vq_launch(embed_bwd_kernel, dim3((B * L + ROW_WARPS - 1) / ROW_WARPS), dim3(ROW_WARPS * 32), 0, stream, ids, B, L, g, S, row0, dtable, drop,
                  vocab);
  VQ_LAUNCH_CHECK();
  return 0;
}

__global__ void shift_right_kernel(const int64_t* __restrict__ labels, int64_t* __restrict__ dec, int B, int T, int start_id, int pad_id) {
  vq_pdl_trigger();
  vq_pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T) return;
  const int t = i % T;
  int64_t v = t == 0 ? (int64_t)start_id : labels[i - 1];
  if (v == -100) v = pad_id;
  dec[i] = v;
}
int shift_right(const int64_t* labels, int64_t* dec_ids, int B, int T, int start_id, int pad_id, cudaStream_t stream) {
  if (B * T <= 0) return 0;
  (void)vq_launch(shift_right_kernel, dim3((B * T + 255) / 256), dim3(256), 0, stream, labels, dec_ids, B, T, start_id, pad_id);
  VQ_LAUNCH_CHECK();
  return 0;
}

__global__ void keymask_kernel(const int64_t* __restrict__ ids, int B, int L, int S, int pad_id, float* __restrict__ enc,
                               float* __restrict__ cross, float* __restrict__ mask01) {
  vq_pdl_trigger();
  vq_pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * (S + 2)) return;
  const int b = i / (S + 2), j = i % (S + 2);
  const bool pad = j < L && ids[b * L + j] == pad_id;
  if (cross) cross[i] = pad ? -1e9f : 0.f;
  if (mask01) mask01[i] = pad ? 0.f : 1.f;
  if (enc && j < S) enc[b * S + j] = pad ? -10000.0f : 0.f;
}
int build_keymasks(const int64_t* ids, int B, int L, int S, int pad_id, float* enc_mask, float* cross_mask, float* mask01,
                   cudaStream_t stream) {
  const int n = B * (S + 2);
  if (n <= 0) return 0;
  (void)vq_launch(keymask_kernel, dim3((n + 255) / 256), dim3(256), 0, stream, ids, B, L, S, pad_id, enc_mask, cross_mask, mask01);
  VQ_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------ VisualEmbedding
VQ_DEVINL void vis_pos5(const float* boxes, int row, float (&p5)[5]) {
  const float4 bx = *reinterpret_cast<const float4*>(boxes + (size_t)row * 4);
  p5[0] = bx.x; p5[1] = bx.y; p5[2] = bx.z; p5[3] = bx.w;
  // get_area as written (modeling_t5_our.py:78-90): (pos[3] - pos[2]) * (pos[1] - pos[0])
  p5[4] = (bx.w - bx.z) * (bx.y - bx.x);
}

__global__ void __launch_bounds__(ROW_WARPS * 32) vis_embed_fwd_kernel(const VisArgs a) {
  vq_pdl_trigger();
  vq_pdl_wait();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (r >= a.B * a.N) return;
  const int b = r / a.N, n = r % a.N;
  float u[RW_CHUNKS][4], t[RW_CHUNKS][4], out[RW_CHUNKS][4];
  // feature branch: RMSNorm(feats Wf^T + bf) * wf
  load_row_f32(u, a.featpre + (size_t)r * DM, lane);
  load_row_f32(t, a.bf, lane);
#pragma unroll
  for (int j = 0; j < RW_CHUNKS; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) u[j][i] += t[j][i];
  float rstd = rsqrtf(row_sumsq(u) / DM + a.eps);
  load_row_f32(t, a.wf, lane);
#pragma unroll
  for (int j = 0; j < RW_CHUNKS; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) out[j][i] = u[j][i] * rstd * t[j][i];
  // position branch: RMSNorm([box, area] Wp^T + bp) * wp
  float p5[5];
  vis_pos5(a.boxes, r, p5);
  load_row_f32(u, a.bp, lane);
#pragma unroll
  for (int j = 0; j < RW_CHUNKS; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float* wrow = a.Wp + (size_t)((lane + 32 * j) * 4 + i) * 5;
      float acc = u[j][i];
#pragma unroll
      for (int k = 0; k < 5; ++k) acc += wrow[k] * p5[k];
      u[j][i] = acc;
    }
  rstd = rsqrtf(row_sumsq(u) / DM + a.eps);
  load_row_f32(t, a.wp, lane);
#pragma unroll
  for (int j = 0; j < RW_CHUNKS; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) out[j][i] += u[j][i] * rstd * t[j][i];
  // + img_order_embedding[0] + shared[V - 1 - n]
  load_row_f32(t, a.img_emb, lane);
  load_row_f32(u, a.shared + (size_t)(a.V - 1 - n) * DM, lane);
#pragma unroll
  for (int j = 0; j < RW_CHUNKS; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) out[j][i] += t[j][i] + u[j][i];
  const size_t orow = (size_t)b * a.S + a.L + n;
  apply_dropout(out, a.drop, orow * DM, lane);
  store_row_f32(a.x + orow * DM, out, lane);
}
int vis_embed_fwd(const VisArgs& a, cudaStream_t stream) {
  const int rows = a.B * a.N;
  if (rows <= 0) return 0;
  (void)vq_launch(vis_embed_fwd_kernel, dim3((rows + ROW_WARPS - 1) / ROW_WARPS), dim3(ROW_WARPS * 32), 0, stream, a);
  VQ_LAUNCH_CHECK();
  return 0;
}

// Column sums over all B*N rows for the 10 per-column gradients (d img_order, d wf, d bf, d wp, d bp, d Wp[:, 0..4]).
// Each warp accumulates into its OWN shared-memory slice with plain read-modify-write (no atomics: a lane owns its
// columns), the CTA folds its slices and writes one partial row to `part[cta][10*768]`; vis_embed_reduce_kernel then sums
// the partials in a fixed order (deterministic) into the gradient arena.
enum { VA_DIMG = 0, VA_DWF, VA_DBF, VA_DWP, VA_DBP, VA_DWP0, VA_COUNT = VA_DWP0 + 5 };
constexpr int VB_WARPS = 4;

// acc4[arr][lane + 32 j] += v: the per-warp accumulators are updated as float4 (one LDS.128 + STS.128 per four columns; the
// scalar form was 264 dependent LDS / FADD / STS chains per lane and row with one warp per scheduler: 135 us for 11 520 rows)
VQ_DEVINL void vb_acc4(float* acc, int arr, int j, int lane, float v0, float v1, float v2, float v3) {
  float4* p4 = reinterpret_cast<float4*>(acc + arr * DM) + lane + 32 * j;
  float4 t = *p4;
  t.x += v0; t.y += v1; t.z += v2; t.w += v3;
  *p4 = t;
}

__global__ void __launch_bounds__(VB_WARPS * 32) vis_embed_bwd_kernel(const VisArgs a, float* __restrict__ part) {
  vq_pdl_trigger();
  vq_pdl_wait();
  extern __shared__ __align__(16) float s_acc[];  // [VB_WARPS][VA_COUNT][DM]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < VB_WARPS * VA_COUNT * DM; i += VB_WARPS * 32) s_acc[i] = 0.f;
  __syncthreads();
  float* acc = s_acc + warp * (VA_COUNT * DM);
  const int rows = a.B * a.N;
  for (int r = blockIdx.x * VB_WARPS + warp; r < rows; r += gridDim.x * VB_WARPS) {
    const int b = r / a.N, n = r % a.N;
    const size_t grow = (size_t)b * a.S + a.L + n;
    float g[RW_CHUNKS][4], u[RW_CHUNKS][4], t[RW_CHUNKS][4];
    load_row_f32(g, a.g + grow * DM, lane);
    apply_dropout(g, a.drop, grow * DM, lane);
    // image-order embedding row 0 and object-order (shared) row V-1-n
    float* dsh = a.dshared + (size_t)(a.V - 1 - n) * DM;
#pragma unroll
    for (int j = 0; j < RW_CHUNKS; ++j) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dsh + (lane + 32 * j) * 4), "f"(g[j][0]), "f"(g[j][1]),
                   "f"(g[j][2]), "f"(g[j][3])
                   : "memory");
      vb_acc4(acc, VA_DIMG, j, lane, g[j][0], g[j][1], g[j][2], g[j][3]);
    }
    // ---- feature branch
    load_row_f32(u, a.featpre + (size_t)r * DM, lane);
    load_row_f32(t, a.bf, lane);
#pragma unroll
    for (int j = 0; j < RW_CHUNKS; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) u[j][i] += t[j][i];
    float rstd = rsqrtf(row_sumsq(u) / DM + a.eps);
    load_row_f32(t, a.wf, lane);
    float dot = 0.f;
    float dx[RW_CHUNKS][4];
#pragma unroll
    for (int j = 0; j < RW_CHUNKS; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        u[j][i] *= rstd;  // uhat
        dx[j][i] = g[j][i] * t[j][i];
        dot += dx[j][i] * u[j][i];
      }
      vb_acc4(acc, VA_DWF, j, lane, g[j][0] * u[j][0], g[j][1] * u[j][1], g[j][2] * u[j][2], g[j][3] * u[j][3]);
    }
    dot = warp_sum(dot) / DM;
#pragma unroll
    for (int j = 0; j < RW_CHUNKS; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) dx[j][i] = rstd * (dx[j][i] - u[j][i] * dot);
      vb_acc4(acc, VA_DBF, j, lane, dx[j][0], dx[j][1], dx[j][2], dx[j][3]);
    }
    store_row_bf16(a.dfeatpre + (size_t)r * DM, dx, lane);
    // ---- position branch
    float p5[5];
    vis_pos5(a.boxes, r, p5);
    load_row_f32(u, a.bp, lane);
#pragma unroll
    for (int j = 0; j < RW_CHUNKS; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float* wrow = a.Wp + (size_t)((lane + 32 * j) * 4 + i) * 5;
        float v = u[j][i];
#pragma unroll
        for (int k = 0; k < 5; ++k) v += wrow[k] * p5[k];
        u[j][i] = v;
      }
    rstd = rsqrtf(row_sumsq(u) / DM + a.eps);
    load_row_f32(t, a.wp, lane);
    dot = 0.f;
#pragma unroll
    for (int j = 0; j < RW_CHUNKS; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        u[j][i] *= rstd;
        dx[j][i] = g[j][i] * t[j][i];
        dot += dx[j][i] * u[j][i];
      }
      vb_acc4(acc, VA_DWP, j, lane, g[j][0] * u[j][0], g[j][1] * u[j][1], g[j][2] * u[j][2], g[j][3] * u[j][3]);
    }
    dot = warp_sum(dot) / DM;
#pragma unroll
    for (int j = 0; j < RW_CHUNKS; ++j) {
      float dv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) dv[i] = rstd * (dx[j][i] - u[j][i] * dot);
      vb_acc4(acc, VA_DBP, j, lane, dv[0], dv[1], dv[2], dv[3]);
#pragma unroll
      for (int k = 0; k < 5; ++k) vb_acc4(acc, VA_DWP0 + k, j, lane, dv[0] * p5[k], dv[1] * p5[k], dv[2] * p5[k], dv[3] * p5[k]);
    }
  }
  __syncthreads();
  float* out = part + (size_t)blockIdx.x * (VA_COUNT * DM);
  for (int i = threadIdx.x; i < VA_COUNT * DM; i += VB_WARPS * 32) {
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < VB_WARPS; ++wv) v += s_acc[wv * (VA_COUNT * DM) + i];
    out[i] = v;
  }
}

__global__ void __launch_bounds__(256) vis_embed_reduce_kernel(const VisArgs a, const float* __restrict__ part, int nparts) {
  vq_pdl_trigger();
  vq_pdl_wait();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= VA_COUNT * DM) return;
  float v = 0.f;
  for (int p = 0; p < nparts; ++p) v += part[(size_t)p * (VA_COUNT * DM) + i];
  const int arr = i / DM, c = i % DM;
  float* dst;
  switch (arr) {
    case VA_DIMG: dst = a.dimg + c; break;
    case VA_DWF: dst = a.dwf + c; break;
    case VA_DBF: dst = a.dbf + c; break;
    case VA_DWP: dst = a.dwp + c; break;
    case VA_DBP: dst = a.dbp + c; break;
    default: dst = a.dWp + (size_t)c * 5 + (arr - VA_DWP0); break;
  }
  *dst += v;
}

int vis_embed_bwd(const VisArgs& a, cudaStream_t stream) {
  const int rows = a.B * a.N;
  if (rows <= 0) return 0;
  VQ_CHECK(a.partials, "vis_embed_bwd: partials scratch buffer missing");
  const int smem = VB_WARPS * VA_COUNT * DM * (int)sizeof(float);
  static bool attr = false;
  if (!attr) {
    VQ_CUDA(cudaFuncSetAttribute(vis_embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  int blocks = (rows + VB_WARPS - 1) / VB_WARPS;
  if (blocks > num_sms()) blocks = num_sms();
  (void)vq_launch(vis_embed_bwd_kernel, dim3(blocks), dim3(VB_WARPS * 32), smem, stream, a, a.partials);
  VQ_LAUNCH_CHECK();
  (void)vq_launch(vis_embed_reduce_kernel, dim3((VA_COUNT * DM + 255) / 256), dim3(256), 0, stream, a, (const float*)a.partials, blocks);
  VQ_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------ device collate
// vqa_data_memory.py:179-187 (box normalisation + clamp) and :386-393 (one-hot labels), for batches assembled from packed
// feature shards: one thread per box / per one-hot element.
__global__ void collate_kernel(const float* __restrict__ boxes_px, const float* __restrict__ wh, int B, int N, float* __restrict__ boxes_out,
                               const int64_t* __restrict__ cate_ids, int n_cate, float* __restrict__ cate_oh,
                               const int64_t* __restrict__ ques_ids, int n_ques, float* __restrict__ ques_oh) {
  vq_pdl_trigger();
  vq_pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * N) {
    const int b = i / N;
    const float w = wh[2 * b], h = wh[2 * b + 1];
    float4 bx = *reinterpret_cast<const float4*>(boxes_px + (size_t)i * 4);
    bx.x = fminf(fmaxf(bx.x / w, 0.f), 1.f);
    bx.y = fminf(fmaxf(bx.y / h, 0.f), 1.f);
    bx.z = fminf(fmaxf(bx.z / w, 0.f), 1.f);
    bx.w = fminf(fmaxf(bx.w / h, 0.f), 1.f);
    *reinterpret_cast<float4*>(boxes_out + (size_t)i * 4) = bx;
  }
  if (cate_ids && i < B * n_cate) cate_oh[i] = cate_ids[i / n_cate] == (int64_t)(i % n_cate) ? 1.f : 0.f;
  if (ques_ids && i < B * n_ques) ques_oh[i] = ques_ids[i / n_ques] == (int64_t)(i % n_ques) ? 1.f : 0.f;
}
int collate_device(const float* boxes_px, const float* wh, int B, int N, float* boxes_out, const int64_t* cate_ids, int n_cate,
                   float* cate_oh, const int64_t* ques_ids, int n_ques, float* ques_oh, cudaStream_t stream) {
  int n = B * N;
  if (cate_ids && B * n_cate > n) n = B * n_cate;
  if (ques_ids && B * n_ques > n) n = B * n_ques;
  if (n <= 0) return 0;
  (void)vq_launch(collate_kernel, dim3((n + 255) / 256), dim3(256), 0, stream, boxes_px, wh, B, N, boxes_out, cate_ids, n_cate, cate_oh, ques_ids,
                  n_ques, ques_oh);
  VQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace vq
