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
  proto_means_kernel<<<B, 192, 0, stream>>>(h, S, split, meanQ, meanV);
  VQ_LAUNCH_CHECK();
  return 0;
}

// one CTA per class: proto[c] = sum_b labels[b,c] * mean[b]  (/ divisor when DIVIDE)
template <bool DIVIDE>
__global__ void __launch_bounds__(192) proto_scatter_kernel(const float* __restrict__ mean, const float* __restrict__ labels, int B,
                                                            int C, float* __restrict__ proto, float* __restrict__ cnt) {
  const int c = blockIdx.x, col = threadIdx.x * 4;
  float4 acc = make_float4(0, 0, 0, 0);
  float n = 0.f;
  for (int b = 0; b < B; ++b) {
    const float l = labels[(size_t)b * C + c];
    n += l;
    if (l != 0.f) {
      const float4 t = *reinterpret_cast<const float4*>(mean + (size_t)b * DM + col);
      acc.x += l * t.x; acc.y += l * t.y; acc.z += l * t.z; acc.w += l * t.w;
    }
  }
  if (DIVIDE) {
    const float d = n <= 0.f ? 1.f : n;  // torch.where(div <= 0, ones, div)
    acc.x /= d; acc.y /= d; acc.z /= d; acc.w /= d;
  }
  *reinterpret_cast<float4*>(proto + (size_t)c * DM + col) = acc;
  if (threadIdx.x == 0) cnt[c] = n;
}
int proto_scatter_mean(const float* mean, const float* labels, int B, int C, float* proto, float* cnt, cudaStream_t stream) {
  proto_scatter_kernel<true><<<C, 192, 0, stream>>>(mean, labels, B, C, proto, cnt);
  VQ_LAUNCH_CHECK();
  return 0;
}
int proto_scatter_sum(const float* mean, const float* labels, int B, int C, float* proto, float* cnt, cudaStream_t stream) {
  proto_scatter_kernel<false><<<C, 192, 0, stream>>>(mean, labels, B, C, proto, cnt);
  VQ_LAUNCH_CHECK();
  return 0;
}
__global__ void __launch_bounds__(192) proto_div_kernel(float* __restrict__ proto, const float* __restrict__ cnt) {
  const int c = blockIdx.x, col = threadIdx.x * 4;
  const float n = cnt[c];
  const float d = n <= 0.f ? 1.f : n;
  float4 t = *reinterpret_cast<float4*>(proto + (size_t)c * DM + col);
  t.x /= d; t.y /= d; t.z /= d; t.w /= d;
  *reinterpret_cast<float4*>(proto + (size_t)c * DM + col) = t;
}
int proto_div(float* proto_sums, const float* cnt, int C, cudaStream_t stream) {
  proto_div_kernel<<<C, 192, 0, stream>>>(proto_sums, cnt);
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
  proto_update_kernel<<<a.CQ + a.CV, 192, 0, stream>>>(a);
  VQ_LAUNCH_CHECK();
  return 0;
}

// cosine_similarity_multi (:434-462): a_n = normalize(tanh(P)), b_n = normalize(tanh(x)) (F.normalize eps 1e-12),
// idx = first argmax_c <a_n[c], b_n>; output = RAW P[idx] written as bf16 into row `out_row` of each batch element's
// [out_pitch_rows, 768] decoder-memory slab (the torch.cat of :615), and optionally as fp32.
// One CTA per batch element; warp w scans classes w, w+8, ...
constexpr int PR_WARPS = 8;
__global__ void __launch_bounds__(PR_WARPS * 32)
proto_retrieve_kernel(const float* __restrict__ P, int C, const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                      int out_pitch_rows, int out_row, int64_t* __restrict__ idx_out, float* __restrict__ out_f32) {
  __shared__ float s_tx[DM];
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
  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int c = warp; c < C; c += PR_WARPS) {
    float dot = 0.f, pn = 0.f;
    for (int k = lane; k < DM; k += 32) {
      const float t = tanhf(P[(size_t)c * DM + k]);
      pn += t * t;
      dot += (t) * (s_tx[k] / xn);
    }
    dot = warp_sum(dot);
    pn = warp_sum(pn);
    const float sim = dot / fmaxf(sqrtf(pn), 1e-12f);
    if (sim > best) { best = sim; besti = c; }  // classes visited in increasing order per warp -> first max
  }
  if (lane == 0) { s_best[warp] = best; s_besti[warp] = besti; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bb = s_best[0];
    int bi = s_besti[0];
    for (int w = 1; w < PR_WARPS; ++w)
      if (s_best[w] > bb || (s_best[w] == bb && s_besti[w] < bi)) { bb = s_best[w]; bi = s_besti[w]; }
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
                   int64_t* idx, float* out_f32, cudaStream_t stream) {
  if (B <= 0) return 0;
  VQ_CHECK(C >= 1, "proto_retrieve: empty bank");
  proto_retrieve_kernel<<<B, PR_WARPS * 32, 0, stream>>>(P, C, x, out, out_pitch_rows, out_row, idx, out_f32);
  VQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace vq
