// Sample-invariant (SI) prototype bank of VQACL (modeling_t5_our.py:434-511, :583-615):
//   proto_means          token mean-pooling of the encoder output (Q = rows [0,20), V = rows [20,S))
//   proto_scatter_*      calculate_current_prototype: per-class scatter-mean + counts           (:500-511)
//   proto_update         update_prototype state machine on explicit device buffers               (:465-498)
//   proto_retrieve       cosine_similarity_multi + feature mix into the decoder memory rows       (:434-462, :615)
// All sums run in a fixed order (deterministic); the argmax is a first-max so exact ties (e.g. all-zero prototype
// rows, whose similarity is exactly 0) resolve to the lowest index like torch.argmax.
#include "ops.h"

namespace vq {

// one CTA per batch element, 192 threads x float4 = 768 columns
__global__ void __launch_bounds__(192) proto_means_kernel(const float* __restrict__ h, int S, int split, float* __restrict__ mq,
                                                          float* __restrict__ mv) {
  vq_pdl_trigger();
  vq_pdl_wait();
  const int b = blockIdx.x, c = threadIdx.x * 4;
  const float* base = h + (size_t)b * S * DM + c;
  float4 aq = make_float4(0, 0, 0, 0), av = make_float4(0, 0, 0, 0);
  const int sp = split < S ? split : S;
  for (int i = 0; i < sp; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(base + (size_t)i * DM);
    aq.x += t.x; aq.y += t.y; aq.z += t.z; aq.w += t.w;
  }
  for (int i = sp; i < S; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(base + (size_t)i * DM);
    av.x += t.x; av.y += t.y; av.z += t.z; av.w += t.w;
  }
  const float nq = (float)sp, nv = (float)(S - sp);
  *reinterpret_cast<float4*>(mq + (size_t)b * DM + c) = make_float4(aq.x / nq, aq.y / nq, aq.z / nq, aq.w / nq);
  *reinterpret_cast<float4*>(mv + (size_t)b * DM + c) = make_float4(av.x / nv, av.y / nv, av.z / nv, av.w / nv);
}
int proto_means(const float* h, int B, int S, int split, float* meanQ, float* meanV, cudaStream_t stream) {
  if (B <= 0) return 0;
  (void)vq_launch(proto_means_kernel, dim3(B), dim3(192), 0, stream, h, S, split, meanQ, meanV);
  VQ_LAUNCH_CHECK();
  return 0;
}

// calculate_current_prototype: proto[c] = sum_b labels[b,c] * mean[b]  (/ divisor when DIVIDE); cnt[c] = sum_b labels[b,c].
// grid (C, 768/128): a CTA owns 128 columns of one class; its 4 warp-groups each reduce a quarter of the batch with
// unconditional FMAs (labels are one-hot weights, so the loads do not depend on them and pipeline 8 deep), then the four
// partials are combined in a fixed order (deterministic; no atomics).
constexpr int PS_COLS = 128, PS_GROUPS = 4;
template <bool DIVIDE>
__global__ void __launch_bounds__(PS_COLS * PS_GROUPS) proto_scatter_kernel(const float* __restrict__ mean, const float* __restrict__ labels, int B,
                                                                            int C, float* __restrict__ proto, float* __restrict__ cnt) {
  vq_pdl_trigger();
  vq_pdl_wait();
  extern __shared__ float ps_smem[];          // wl[B] | part[PS_GROUPS][PS_COLS]
  float* wl = ps_smem;
  float* part = ps_smem + ((B + 3) & ~3);
  const int c = blockIdx.x, col = blockIdx.y * PS_COLS + (threadIdx.x & (PS_COLS - 1)), grp = threadIdx.x / PS_COLS;
  for (int b = threadIdx.x; b < B; b += PS_COLS * PS_GROUPS) wl[b] = labels[(size_t)b * C + c];
  __syncthreads();
  const int per = (B + PS_GROUPS - 1) / PS_GROUPS;
  const int b0 = grp * per, b1 = min(B, b0 + per);
  float acc = 0.f;
  int b = b0;
  for (; b + 8 <= b1; b += 8) {
    float m[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) m[u] = mean[(size_t)(b + u) * DM + col];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += wl[b + u] * m[u];
  }
  for (; b < b1; ++b) acc += wl[b] * mean[(size_t)b * DM + col];
  part[grp * PS_COLS + (threadIdx.x & (PS_COLS - 1))] = acc;
  __syncthreads();
  if (grp == 0) {
    float n = 0.f;
    for (int i = 0; i < B; ++i) n += wl[i];            // exact: one-hot labels, integer-valued sums
    float v = part[threadIdx.x] + part[PS_COLS + threadIdx.x] + part[2 * PS_COLS + threadIdx.x] + part[3 * PS_COLS + threadIdx.x];
    if (DIVIDE) v /= (n <= 0.f ? 1.f : n);             // torch.where(div <= 0, ones, div)
    proto[(size_t)c * DM + col] = v;
    if (blockIdx.y == 0 && threadIdx.x == 0) cnt[c] = n;
  }
}
template <bool DIVIDE>
static int proto_scatter_launch(const float* mean, const float* labels, int B, int C, float* proto, float* cnt, cudaStream_t stream) {
  if (C <= 0) return 0;
  const size_t smem = (((size_t)B + 3) & ~(size_t)3) * 4 + PS_GROUPS * PS_COLS * 4;
  VQ_CHECK(smem <= 48 * 1024, "proto_scatter: batch %d too large for the label staging buffer", B);
  (void)vq_launch(proto_scatter_kernel<DIVIDE>, dim3(C, DM / PS_COLS), dim3(PS_COLS * PS_GROUPS), smem, stream, mean, labels, B, C, proto, cnt);
  VQ_LAUNCH_CHECK();
  return 0;
}
int proto_scatter_mean(const float* mean, const float* labels, int B, int C, float* proto, float* cnt, cudaStream_t stream) {
  return proto_scatter_launch<true>(mean, labels, B, C, proto, cnt, stream);
}
int proto_scatter_sum(const float* mean, const float* labels, int B, int C, float* proto, float* cnt, cudaStream_t stream) {
  return proto_scatter_launch<false>(mean, labels, B, C, proto, cnt, stream);
}
__global__ void __launch_bounds__(192) proto_div_kernel(float* __restrict__ proto, const float* __restrict__ cnt) {
  vq_pdl_trigger();
  vq_pdl_wait();
  const int c = blockIdx.x, col = threadIdx.x * 4;
  const float n = cnt[c];
  const float d = n <= 0.f ? 1.f : n;
  float4 t = *reinterpret_cast<float4*>(proto + (size_t)c * DM + col);
  t.x /= d; t.y /= d; t.z /= d; t.w /= d;
  *reinterpret_cast<float4*>(proto + (size_t)c * DM + col) = t;
}
int proto_div(float* proto_sums, const float* cnt, int C, cudaStream_t stream) {
  (void)vq_launch(proto_div_kernel, dim3(C), dim3(192), 0, stream, proto_sums, cnt);
  VQ_LAUNCH_CHECK();
  return 0;
}

// update_prototype (modeling_t5_our.py:465-498). Qproto is ONE persistent buffer: in the reference, for task t > 0,
// `Q_prototype` shares storage with `Q_task_mem_proto[t]` (:490) and row t is overwritten in place (:491), so the bank
// and the task memory are the same tensor; for t == 0 the bank is the current batch's class means.
//   first step of task t (t not in Q_task_cur_proto):  V = curV; nums = counts;
//         t == 0: Q = curQ;                 t > 0: Q[t] = curQ[t], other rows keep the bank inherited from task t-1
//   later steps: t == 0: Q = curQ (no EMA)
//         t > 0: rows c != t: has_mem ? alpha*Q[c] + (1-alpha)*curQ[c] : curQ[c];   row t: curQ[t]
//         V = beta*V + (1-beta)*curV;  nums += counts
// blocks [0,CQ) handle Q rows, [CQ, CQ+CV) handle V rows.
__global__ void __launch_bounds__(192) proto_update_kernel(const ProtoUpdateArgs a) {
  vq_pdl_trigger();
  vq_pdl_wait();
  const int col = threadIdx.x * 4;
  if ((int)blockIdx.x < a.CQ) {
    const int c = blockIdx.x;
    const float4 cur = *reinterpret_cast<const float4*>(a.curQ + (size_t)c * DM + col);
    float4* dst = reinterpret_cast<float4*>(a.Qproto + (size_t)c * DM + col);
    if (a.first_step_of_task) {
      if (a.task_id == 0 || c == a.task_id) *dst = cur;
    } else if (a.task_id == 0 || c == a.task_id || !a.has_mem) {
      *dst = cur;
    } else {
      const float4 old = *dst;
      const float al = a.alpha, be = 1.f - a.alpha;
      *dst = make_float4(al * old.x + be * cur.x, al * old.y + be * cur.y, al * old.z + be * cur.z, al * old.w + be * cur.w);
    }
    if (threadIdx.x == 0) a.numQ[c] = a.first_step_of_task ? a.cntQ[c] : a.numQ[c] + a.cntQ[c];
  } else {
    const int c = blockIdx.x - a.CQ;
    const float4 cur = *reinterpret_cast<const float4*>(a.curV + (size_t)c * DM + col);
    float4* dst = reinterpret_cast<float4*>(a.Vproto + (size_t)c * DM + col);
    if (a.first_step_of_task) {
      *dst = cur;
    } else {
      const float4 old = *dst;
      const float be = a.beta, om = 1.f - a.beta;
      *dst = make_float4(be * old.x + om * cur.x, be * old.y + om * cur.y, be * old.z + om * cur.z, be * old.w + om * cur.w);
    }
    if (threadIdx.x == 0) a.numV[c] = a.first_step_of_task ? a.cntV[c] : a.numV[c] + a.cntV[c];
  }
}
int proto_update(const ProtoUpdateArgs& a, cudaStream_t stream) {
  VQ_CHECK(a.task_id >= 0 && a.task_id < a.CQ, "proto_update: task id %d outside [0,%d)", a.task_id, a.CQ);
  (void)vq_launch(proto_update_kernel, dim3(a.CQ + a.CV), dim3(192), 0, stream, a);
  VQ_LAUNCH_CHECK();
  return 0;
}

// cosine_similarity_multi (:434-462): a_n = normalize(tanh(P)), b_n = normalize(tanh(x)) (F.normalize eps 1e-12),
// idx = first argmax_c <a_n[c], b_n>; output = RAW P[idx] written as bf16 into row `out_row` of each batch element's
// [out_pitch_rows, 768] decoder-memory slab (the torch.cat of :615), and optionally as fp32.
// Step 1 (proto_normalize_kernel, one CTA per class): Pn[c] = tanh(P[c]) / max(||tanh(P[c])||, 1e-12), once per call
// instead of once per batch element. Step 2 (proto_retrieve_kernel, one CTA per batch element): warp w scans classes
// w, w+8, ... with warp-shuffle dot products.
constexpr int PR_WARPS = 8;
__global__ void __launch_bounds__(PR_WARPS * 32) proto_normalize_kernel(const float* __restrict__ P, float* __restrict__ Pn) {
  vq_pdl_trigger();
  vq_pdl_wait();
  __shared__ float s_red[PR_WARPS];
  const int c = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float t[DM / (PR_WARPS * 32)];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < DM / (PR_WARPS * 32); ++j) {
    t[j] = tanhf(P[(size_t)c * DM + threadIdx.x + j * PR_WARPS * 32]);
    ss += t[j] * t[j];
  }
  ss = warp_sum(ss);
  if (lane == 0) s_red[warp] = ss;
  __syncthreads();
  float n = 0.f;
#pragma unroll
  for (int w = 0; w < PR_WARPS; ++w) n += s_red[w];
  n = fmaxf(sqrtf(n), 1e-12f);
#pragma unroll
  for (int j = 0; j < DM / (PR_WARPS * 32); ++j) Pn[(size_t)c * DM + threadIdx.x + j * PR_WARPS * 32] = t[j] / n;
}

__global__ void __launch_bounds__(PR_WARPS * 32)
proto_retrieve_kernel(const float* __restrict__ P, const float* __restrict__ Pn, int C, const float* __restrict__ x,
                      __nv_bfloat16* __restrict__ out, int out_pitch_rows, int out_row, int64_t* __restrict__ idx_out,
                      float* __restrict__ out_f32) {
  vq_pdl_trigger();
  vq_pdl_wait();
  __shared__ __align__(16) float s_tx[DM];
  __shared__ float s_best[PR_WARPS];
  __shared__ int s_besti[PR_WARPS];
  __shared__ float s_red[PR_WARPS];
  __shared__ int s_idx;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // tanh(x) and its norm
  float ss = 0.f;
  for (int c = threadIdx.x; c < DM; c += PR_WARPS * 32) {
    const float t = tanhf(x[(size_t)b * DM + c]);
    s_tx[c] = t;
    ss += t * t;
  }
  ss = warp_sum(ss);
  if (lane == 0) s_red[warp] = ss;
  __syncthreads();
  float xn = 0.f;
#pragma unroll
  for (int w = 0; w < PR_WARPS; ++w) xn += s_red[w];
  xn = fmaxf(sqrtf(xn), 1e-12f);
  for (int c = threadIdx.x; c < DM; c += PR_WARPS * 32) s_tx[c] /= xn;   // each thread rescales the elements it wrote
  __syncthreads();
  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int c = warp; c < C; c += PR_WARPS) {
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < DM / 128; ++j) {
      const int k = (lane + 32 * j) * 4;
      const float4 a = *reinterpret_cast<const float4*>(Pn + (size_t)c * DM + k);
      const float4 t = *reinterpret_cast<const float4*>(s_tx + k);
      dot += a.x * t.x + a.y * t.y + a.z * t.z + a.w * t.w;
    }
    dot = warp_sum(dot);
    if (dot > best) { best = dot; besti = c; }  // classes visited in increasing order per warp -> first max
  }
  if (lane == 0) { s_best[warp] = best; s_besti[warp] = besti; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bb = s_best[0];
    int bi = s_besti[0];
    for (int w = 1; w < PR_WARPS; ++w)
      if (s_best[w] > bb || (s_best[w] == bb && s_besti[w] < bi)) { bb = s_best[w]; bi = s_besti[w]; }
    if (bi < 0 || bi >= C) bi = 0;   // all similarities NaN: keep the access in range
    s_idx = bi;
    idx_out[b] = bi;
  }
  __syncthreads();
  const int bi = s_idx;
  for (int c = threadIdx.x; c < DM; c += PR_WARPS * 32) {
    const float v = P[(size_t)bi * DM + c];
    if (out) out[((size_t)b * out_pitch_rows + out_row) * DM + c] = __float2bfloat16_rn(v);
    if (out_f32) out_f32[(size_t)b * DM + c] = v;
  }
}
int proto_retrieve(const float* P, int C, const float* x, int B, __nv_bfloat16* out, int out_pitch_rows, int out_row,
                   int64_t* idx, float* out_f32, float* scratch, cudaStream_t stream) {
  if (B <= 0) return 0;
  VQ_CHECK(C >= 1, "proto_retrieve: empty bank");
  VQ_CHECK(scratch, "proto_retrieve: scratch [C,768] fp32 required");
  (void)vq_launch(proto_normalize_kernel, dim3(C), dim3(PR_WARPS * 32), 0, stream, P, scratch);
  VQ_LAUNCH_CHECK();
  (void)vq_launch(proto_retrieve_kernel, dim3(B), dim3(PR_WARPS * 32), 0, stream, P, (const float*)scratch, C, x, out, out_pitch_rows, out_row, idx,
                  out_f32);
  VQ_LAUNCH_CHECK();
  return 0;
}

// memory_loss (VL-T5/nextqa/modeling_t5_nextqa.py:544-555, called at VL-T5/src/modeling_t5_our.py:591 BEFORE the bank
// update): the prototype pull loss  loss = mean_b sum_d (mean_t(h)[b,d] - (labels @ P.detach())[b,d])^2  for the Q side
// (question-type bank) and the V side (object-category bank). One CTA per (sample, side): diff[b] = mean[b] - sum_c
// labels[b,c] P[c] is kept (fp32) for the backward pass, its squared norm goes to ssq[side][b]; a second single-CTA kernel
// averages the B values in a fixed order (deterministic) into loss[side].
__global__ void __launch_bounds__(192) proto_memloss_kernel(const float* __restrict__ meanQ, const float* __restrict__ meanV,
                                                            const float* __restrict__ lq, const float* __restrict__ lv,
                                                            const float* __restrict__ PQ, const float* __restrict__ PV, int CQ, int CV,
                                                            int B, float* __restrict__ diffQ, float* __restrict__ diffV,
                                                            float* __restrict__ ssq) {
  vq_pdl_trigger();
  vq_pdl_wait();
  __shared__ float s_w[6];
  const int b = blockIdx.x, side = blockIdx.y, col = threadIdx.x * 4;
  const float* mean = side ? meanV : meanQ;
  const float* lab = (side ? lv : lq) + (size_t)b * (side ? CV : CQ);
  const float* P = side ? PV : PQ;
  const int C = side ? CV : CQ;
  float4 acc = *reinterpret_cast<const float4*>(mean + (size_t)b * DM + col);
  for (int c = 0; c < C; ++c) {
    const float w = lab[c];
    if (w != 0.f) {                       // one-hot labels: one row of the bank per sample
      const float4 p = *reinterpret_cast<const float4*>(P + (size_t)c * DM + col);
      acc.x -= w * p.x; acc.y -= w * p.y; acc.z -= w * p.z; acc.w -= w * p.w;
    }
  }
  *reinterpret_cast<float4*>((side ? diffV : diffQ) + (size_t)b * DM + col) = acc;
  float sq = warp_sum(acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) ssq[(size_t)side * B + b] = s_w[0] + s_w[1] + s_w[2] + s_w[3] + s_w[4] + s_w[5];
}
__global__ void __launch_bounds__(64) proto_memloss_final_kernel(const float* __restrict__ ssq, int B, float* __restrict__ loss2) {
  vq_pdl_trigger();
  vq_pdl_wait();
  const int side = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int b = lane; b < B; b += 32) acc += ssq[(size_t)side * B + b];
  acc = warp_sum(acc);
  if (lane == 0) loss2[side] = acc / (float)B;
}
int proto_memory_loss(const float* meanQ, const float* meanV, const float* ques_labels, const float* cate_labels, const float* PQ,
                      const float* PV, int CQ, int CV, int B, float* diffQ, float* diffV, float* ssq, float* loss2, cudaStream_t stream) {
  if (B <= 0) return 0;
  (void)vq_launch(proto_memloss_kernel, dim3(B, 2), dim3(192), 0, stream, meanQ, meanV, ques_labels, cate_labels, PQ, PV, CQ, CV, B, diffQ,
                  diffV, ssq);
  VQ_LAUNCH_CHECK();
  (void)vq_launch(proto_memloss_final_kernel, dim3(1), dim3(64), 0, stream, (const float*)ssq, B, loss2);
  VQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace vq
