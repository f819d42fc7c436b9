// LM-head cross-entropy over the 32 200-token vocabulary (modeling_t5_our.py:666-686; vqa_model.py:46-54).
// The logits come out of the tcgen05 GEMM as bf16 [M, V] (pitch ld). One CTA owns one row:
//   ce_fwd : single pass online log-sum-exp (fp32), loss = lse - logit[label], 0 for ignore_index -100
//   ce_bwd : in place logits <- (softmax - onehot) * w[row]  — the bf16 A operand of the dX / dE GEMMs
//   loss_tail: the per-sample masked mean x soft score x batch mean of VLT5VQA.train_step, plus dL/dloss_row
//   argmax_rows: greedy token choice (first max) for generation
#include "ops.h"

namespace vq {

constexpr int CE_THREADS = 256;

VQ_DEVINL void lse_combine(float& m, float& s, float m2, float s2) {
  const float mm = fmaxf(m, m2);
  if (mm == -INFINITY) { m = mm; s = 0.f; return; }
  s = s * __expf(m - mm) + s2 * __expf(m2 - mm);
  m = mm;
}

__global__ void __launch_bounds__(CE_THREADS)
ce_fwd_kernel(const __nv_bfloat16* __restrict__ logits, int ld, int V, const int64_t* __restrict__ labels, float* __restrict__ lse,
              float* __restrict__ loss) {
  vq_pdl_trigger();
  vq_pdl_wait();
  __shared__ float s_m[CE_THREADS / 32], s_s[CE_THREADS / 32];
  const int r = blockIdx.x;
  const __nv_bfloat16* row = logits + (size_t)r * ld;
  float m = -INFINITY, s = 0.f;
  const int nv = V / 8;
  for (int i = threadIdx.x; i < nv; i += CE_THREADS) {
    const uint4 u = reinterpret_cast<const uint4*>(row)[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    float f[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = unpack_bf16(w[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
    float mx = f[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) mx = fmaxf(mx, f[j]);
    float ls = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) ls += __expf(f[j] - mx);
    lse_combine(m, s, mx, ls);
  }
  for (int i = nv * 8 + threadIdx.x; i < V; i += CE_THREADS) lse_combine(m, s, __bfloat162float(row[i]), 1.f);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    lse_combine(m, s, m2, s2);
  }
  if ((threadIdx.x & 31) == 0) { s_m[threadIdx.x >> 5] = m; s_s[threadIdx.x >> 5] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < CE_THREADS / 32; ++w) lse_combine(m, s, s_m[w], s_s[w]);
    const float l = m + logf(s);
    lse[r] = l;
    const int64_t lab = labels[r];
    loss[r] = (lab >= 0 && lab < V) ? l - __bfloat162float(row[lab]) : 0.f;
  }
}
int ce_fwd(const __nv_bfloat16* logits, int ld, int M, int V, const int64_t* labels, float* lse, float* loss, cudaStream_t stream) {
  if (M <= 0) return 0;
  VQ_CHECK(ld % 8 == 0, "ce_fwd: logits pitch %d must be a multiple of 8", ld);
  (void)vq_launch(ce_fwd_kernel, dim3(M), dim3(CE_THREADS), 0, stream, logits, ld, V, labels, lse, loss);
  VQ_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(CE_THREADS)
ce_bwd_kernel(__nv_bfloat16* __restrict__ logits, int ld, int V, const int64_t* __restrict__ labels, const float* __restrict__ lse,
              const float* __restrict__ w, const float* __restrict__ gscale) {
  vq_pdl_trigger();
  vq_pdl_wait();
  const int r = blockIdx.x;
  __nv_bfloat16* row = logits + (size_t)r * ld;
  const int64_t lab = labels[r];
  const float wr = (lab >= 0 && lab < V) ? w[r] * (gscale ? *gscale : 1.f) : 0.f;   // gscale = upstream d(loss) (autograd)
  const float l = lse[r];
  const int nv = ld / 8;  // also clears the pitch padding beyond V
  for (int i = threadIdx.x; i < nv; i += CE_THREADS) {
    uint4 u = reinterpret_cast<uint4*>(row)[i];
    uint32_t ww[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = unpack_bf16(ww[j]);
      const int c = i * 8 + 2 * j;
      float g0 = 0.f, g1 = 0.f;
      if (wr != 0.f) {
        g0 = c < V ? (__expf(t.x - l) - (c == lab ? 1.f : 0.f)) * wr : 0.f;
        g1 = c + 1 < V ? (__expf(t.y - l) - (c + 1 == lab ? 1.f : 0.f)) * wr : 0.f;
      }
      ww[j] = pack_bf16(g0, g1);
    }
    reinterpret_cast<uint4*>(row)[i] = make_uint4(ww[0], ww[1], ww[2], ww[3]);
  }
}
int ce_bwd(__nv_bfloat16* logits, int ld, int M, int V, const int64_t* labels, const float* lse, const float* w, const float* gscale,
           cudaStream_t stream) {
  if (M <= 0) return 0;
  VQ_CHECK(ld % 8 == 0, "ce_bwd: logits pitch %d must be a multiple of 8", ld);
  (void)vq_launch(ce_bwd_kernel, dim3(M), dim3(CE_THREADS), 0, stream, logits, ld, V, labels, lse, w, gscale);
  VQ_LAUNCH_CHECK();
  return 0;
}

// one CTA: B*T is small (<= a few thousand)
__global__ void __launch_bounds__(256) loss_tail_kernel(const float* __restrict__ loss_rows, const int64_t* __restrict__ labels,
                                                        const float* __restrict__ scores, int B, int T, float* __restrict__ loss_out,
                                                        float* __restrict__ w_rows) {
  vq_pdl_trigger();
  vq_pdl_wait();
  __shared__ float s_part[256];
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) {
    float n = 0.f, s = 0.f;
    for (int t = 0; t < T; ++t) {
      const bool valid = labels[b * T + t] != -100;
      if (valid) { n += 1.f; s += loss_rows[b * T + t]; }
    }
    const float d = fmaxf(n, 1.f);
    acc += s / d * scores[b];
    if (w_rows)
      for (int t = 0; t < T; ++t) w_rows[b * T + t] = labels[b * T + t] != -100 ? scores[b] / (d * (float)B) : 0.f;
  }
  s_part[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_part[threadIdx.x] += s_part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss_out = s_part[0] / (float)B;
}
int loss_tail(const float* loss_rows, const int64_t* labels, const float* scores, int B, int T, float* loss_out, float* w_rows,
              cudaStream_t stream) {
  (void)vq_launch(loss_tail_kernel, dim3(1), dim3(256), 0, stream, loss_rows, labels, scores, B, T, loss_out, w_rows);
  VQ_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(CE_THREADS)
argmax_rows_kernel(const __nv_bfloat16* __restrict__ logits, int ld, int V, int64_t* __restrict__ out) {
  vq_pdl_trigger();
  vq_pdl_wait();
  __shared__ float s_v[CE_THREADS / 32];
  __shared__ int s_i[CE_THREADS / 32];
  const int r = blockIdx.x;
  const __nv_bfloat16* row = logits + (size_t)r * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += CE_THREADS) {
    const float v = __bfloat162float(row[i]);
    if (v > best) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float v2 = __shfl_xor_sync(0xffffffffu, best, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
    if (v2 > best || (v2 == best && i2 < bi)) { best = v2; bi = i2; }
  }
  if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < CE_THREADS / 32; ++w)
      if (s_v[w] > best || (s_v[w] == best && s_i[w] < bi)) { best = s_v[w]; bi = s_i[w]; }
    out[r] = bi;
  }
}
int argmax_rows(const __nv_bfloat16* logits, int ld, int M, int V, int64_t* out, cudaStream_t stream) {
  if (M <= 0) return 0;
  (void)vq_launch(argmax_rows_kernel, dim3(M), dim3(CE_THREADS), 0, stream, logits, ld, V, out);
  VQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace vq
