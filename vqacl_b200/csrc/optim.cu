// Step tail of Trainer.train_step (vqacl.py:461-487): clip_grad_norm_(5.0) + HF-4.2.1 AdamW over the flat parameter
// arena, fused with the bf16 refresh of the GEMM weight copies. HBM-bound: per element it reads p, g, m, v (16 B) and
// writes p, m, v, bf16(p) (14 B).
//   HF AdamW (transformers 4.2.1 optimization.py, correct_bias=True), per element:
//     m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps);
//     then p -= lr * wd * p          (decay after the step, with lr, on the decayed group only)
#include "ops.h"

namespace vq {

constexpr int SQ_THREADS = 256;
constexpr int SQ_MAX_BLOCKS = 1184;  // 148 * 8

__global__ void __launch_bounds__(SQ_THREADS) sumsq_partial_kernel(const float* __restrict__ g, size_t n4, size_t n, float* __restrict__ partials) {
  vq_pdl_trigger();
  vq_pdl_wait();
  __shared__ float s_w[SQ_THREADS / 32];
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)SQ_THREADS + threadIdx.x; i < n4; i += (size_t)gridDim.x * SQ_THREADS) {
    const float4 t = reinterpret_cast<const float4*>(g)[i];
    acc += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
  }
  if (blockIdx.x == 0)
    for (size_t i = n4 * 4 + threadIdx.x; i < n; i += SQ_THREADS) acc += g[i] * g[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < SQ_THREADS / 32; ++w) s += s_w[w];
    partials[blockIdx.x] = s;
  }
}
__global__ void __launch_bounds__(256) sumsq_final_kernel(const float* __restrict__ partials, int n, float* __restrict__ out) {
  vq_pdl_trigger();
  vq_pdl_wait();
  __shared__ double s_p[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += (double)partials[i];
  s_p[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_p[threadIdx.x] += s_p[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)s_p[0];
}
int grad_sumsq(const float* g, size_t n, float* partials, float* out, cudaStream_t stream) {
  VQ_CHECK((reinterpret_cast<uintptr_t>(g) & 15) == 0, "grad_sumsq: gradient arena must be 16-byte aligned");
  const size_t n4 = n / 4;
  size_t want = (n4 + SQ_THREADS - 1) / SQ_THREADS;
  const int blocks = (int)(want < 1 ? 1 : (want > SQ_MAX_BLOCKS ? SQ_MAX_BLOCKS : want));
  (void)vq_launch(sumsq_partial_kernel, dim3(blocks), dim3(SQ_THREADS), 0, stream, g, n4, n, partials);
  VQ_LAUNCH_CHECK();
  (void)vq_launch(sumsq_final_kernel, dim3(1), dim3(256), 0, stream, partials, blocks, out);
  VQ_LAUNCH_CHECK();
  return 0;
}

// sum of squares over several disjoint ranges of one arena (the slices a rank owns after a bucketed reduce-scatter):
// one partial launch per range into consecutive segments of `partials`, one final reduction over all of them
int grad_sumsq_ranges(const float* g, const int64_t* begin, const int64_t* end, int n_ranges, float* partials, int partials_cap,
                      float* out, cudaStream_t stream) {
  int used = 0;
  for (int i = 0; i < n_ranges; ++i) {
    const int64_t n = end[i] - begin[i];
    if (n <= 0) continue;
    const float* gi = g + begin[i];
    VQ_CHECK((reinterpret_cast<uintptr_t>(gi) & 15) == 0, "grad_sumsq_ranges: range %d is not 16-byte aligned", i);
    const size_t n4 = (size_t)n / 4;
    size_t want = (n4 + SQ_THREADS - 1) / SQ_THREADS;
    int blocks = (int)(want < 1 ? 1 : (want > 296 ? 296 : want));
    VQ_CHECK(used + blocks <= partials_cap, "grad_sumsq_ranges: %d ranges need more than %d partials", n_ranges, partials_cap);
    (void)vq_launch(sumsq_partial_kernel, dim3(blocks), dim3(SQ_THREADS), 0, stream, gi, n4, (size_t)n, partials + used);
    VQ_LAUNCH_CHECK();
    used += blocks;
  }
  (void)vq_launch(sumsq_final_kernel, dim3(1), dim3(256), 0, stream, (const float*)partials, used, out);
  VQ_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(256) adamw_kernel(const AdamArgs a, float step_size, float clip_max) {
  vq_pdl_trigger();
  vq_pdl_wait();
  float coef = 1.f;
  if (a.sumsq && clip_max > 0.f) {
    const float c = clip_max / (sqrtf(*a.sumsq) + 1e-6f);  // torch clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6)
    coef = c < 1.f ? c : 1.f;
  }
  const float b1 = a.beta1, b2 = a.beta2;
  const size_t n4 = a.n / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  // two float4 groups per thread and iteration: all eight 16-byte loads are issued before the first use, so a small grid
  // (the overlapped mode runs 2 CTAs per SM next to a persistent GEMM CTA) still keeps enough bytes in flight
  for (size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i0 < n4; i0 += 2 * stride) {
    const size_t idx[2] = {i0, i0 + stride};
    float4 p[2], g4[2], m[2], v[2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (idx[u] < n4) {
        p[u] = reinterpret_cast<float4*>(a.p)[idx[u]];
        g4[u] = reinterpret_cast<const float4*>(a.g)[idx[u]];
        m[u] = reinterpret_cast<float4*>(a.m)[idx[u]];
        v[u] = reinterpret_cast<float4*>(a.v)[idx[u]];
      }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (idx[u] < n4) {
        const size_t i = idx[u];
        const float wd = (i * 4 < a.n_decay) ? a.lr * a.weight_decay : 0.f;  // n_decay is a multiple of 4 (host-checked)
        float* pp = &p[u].x; const float* gg = &g4[u].x; float* mm = &m[u].x; float* vv = &v[u].x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float g = gg[k] * coef;
          mm[k] = b1 * mm[k] + (1.f - b1) * g;
          vv[k] = b2 * vv[k] + (1.f - b2) * g * g;
          float x = pp[k] - step_size * (mm[k] / (sqrtf(vv[k]) + a.eps));
          x = x - wd * x;
          pp[k] = x;
        }
        reinterpret_cast<float4*>(a.p)[i] = p[u];
        reinterpret_cast<float4*>(a.m)[i] = m[u];
        reinterpret_cast<float4*>(a.v)[i] = v[u];
        if (a.p_bf16) reinterpret_cast<uint2*>(a.p_bf16)[i] = make_uint2(pack_bf16(p[u].x, p[u].y), pack_bf16(p[u].z, p[u].w));
      }
  }
}
int adamw_hf(const AdamArgs& a, cudaStream_t stream) {
  VQ_CHECK(a.n % 4 == 0 && a.n_decay % 4 == 0, "adamw: arena sizes must be multiples of 4 (n=%zu n_decay=%zu)", a.n, a.n_decay);
  VQ_CHECK(a.step >= 1, "adamw: step must be >= 1");
  if (a.n == 0) return 0;
  const double bc1 = 1.0 - pow((double)a.beta1, (double)a.step);
  const double bc2 = 1.0 - pow((double)a.beta2, (double)a.step);
  const float step_size = (float)((double)a.lr * sqrt(bc2) / bc1);
  const size_t n4 = a.n / 4;
  size_t want = (n4 + 255) / 256;
  size_t cap = (size_t)num_sms() * 8;
  if (a.max_blocks > 0 && (size_t)a.max_blocks < cap) cap = (size_t)a.max_blocks;
  want = (want + 1) / 2;
  const int blocks = (int)(want > cap ? cap : (want < 1 ? 1 : want));
  (void)vq_launch(adamw_kernel, dim3(blocks), dim3(256), 0, stream, a, step_size, a.max_norm);
  VQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace vq
