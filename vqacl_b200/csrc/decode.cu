// Greedy answer generation (VLT5VQA.test_step -> HF 4.2.1 generate/greedy_search; vqa_model.py:112-116,
// modeling_t5_our.py:607-615,715-735): encoder once, retrieval from the frozen prototype banks
// once (it is step-invariant), cross-attention K/V of all decoder layers once, then one KV-cached decoder step per
// token with M = B rows. Per step and layer the self-attention q|k|v row of the new token is written by ONE GEMM straight
// into the [B, max_len, 3d] cache at position t, and attention reads the cache in place.
#include "engine.h"

#include <math.h>

namespace vq {

struct DecodeWs {
  bf16* cache;        // [Ld][B, max_len, 3d]  q|k|v of every generated position
  float* pval;        // [B, slots] per-tile maxima of the LM-head GEMM (EPI_ARGMAX): the [B, 32200] logits are never materialised
  int* pidx;          // [B, slots] their column indices
  int slots;          // row pitch of pval / pidx (>= 2 * ceil(V / 256), multiple of 8)
  int64_t* cur;       // [B] current input token
  int* unfinished;    // [B]
  int* n_unfinished;  // [max_len] rows still unfinished after each generated column
};

static int64_t decode_carve(const Engine& e, uint8_t* base, int B, int max_len, DecodeWs* out) {
  const int d = e.cfg.d_model, Ld = e.cfg.n_dec_layers;
  const int slots = (((e.cfg.vocab_size + 255) / 256) * 2 + 7) & ~7;
  int64_t off = 0;
  auto take = [&](size_t bytes) {
    off = (off + 255) / 256 * 256;
    uint8_t* p = base ? base + off : nullptr;
    off += (int64_t)bytes;
    return p;
  };
  DecodeWs w;
  w.cache = reinterpret_cast<bf16*>(take((size_t)Ld * B * max_len * 3 * d * sizeof(bf16)));
  w.pval = reinterpret_cast<float*>(take((size_t)B * slots * sizeof(float)));
  w.pidx = reinterpret_cast<int*>(take((size_t)B * slots * sizeof(int)));
  w.slots = slots;
  w.cur = reinterpret_cast<int64_t*>(take((size_t)B * sizeof(int64_t)));
  w.unfinished = reinterpret_cast<int*>(take((size_t)B * sizeof(int)));
  w.n_unfinished = reinterpret_cast<int*>(take((size_t)(max_len + 1) * sizeof(int)));
  if (out) *out = w;
  return (off + 255) / 256 * 256;
}

__global__ void decode_init_kernel(int64_t* __restrict__ out, int max_len, int64_t* __restrict__ cur, int* __restrict__ unfinished,
                                   int B, int start_id) {
  vq_pdl_trigger();
  vq_pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  out[(size_t)b * max_len] = start_id;
  cur[b] = start_id;
  unfinished[b] = 1;
}

// Finish the fused LM-head argmax (merge the per-tile maxima: largest value, lowest index among equals = torch.argmax's first
// maximum) and do HF greedy_search's bookkeeping: next = argmax * unfinished + pad * (1 - unfinished);
// unfinished &= (next != eos). One warp per row.
__global__ void __launch_bounds__(256)
decode_advance_kernel(const float* __restrict__ pval, const int* __restrict__ pidx, int slots, int nslots, int64_t* __restrict__ out,
                      int max_len, int col, int64_t* __restrict__ cur, int* __restrict__ unfinished, int* __restrict__ n_unfinished,
                      int B, int pad_id, int eos_id) {
  vq_pdl_trigger();
  vq_pdl_wait();
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = lane; i < nslots; i += 32) {
    const float v = pval[(size_t)b * slots + i];
    const int id = pidx[(size_t)b * slots + i];
    if (v > best || (v == best && id < bi)) { best = v; bi = id; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float v2 = __shfl_xor_sync(0xffffffffu, best, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
    if (v2 > best || (v2 == best && i2 < bi)) { best = v2; bi = i2; }
  }
  if (lane != 0) return;
  const int u = unfinished[b];
  const int64_t tok = u ? (int64_t)bi : (int64_t)pad_id;
  out[(size_t)b * max_len + col] = tok;
  cur[b] = tok;
  const int nu = u && tok != eos_id;
  unfinished[b] = nu;
  if (nu) atomicAdd(&n_unfinished[col], 1);
}

static int decode_step(Engine& e, const DecodeWs& dw, int B, int S2, int max_len, int t, cudaStream_t st) {
  const vqacl_config& c = e.cfg;
  Workspace& w = e.w;
  const int d = c.d_model, f = c.d_ff, H = c.n_heads, Ld = c.n_dec_layers;
  const int ldkv = Ld * 2 * d;
  const long long cache_bs = (long long)max_len * 3 * d;
  VQ_TRY(embed_fwd(dw.cur, B, 1, e.P + e.o_shared, w.y[0], 1, 0, Dropout(), c.vocab_size, nullptr, st));
  for (int l = 0; l < Ld; ++l) {
    const DecLayer& P = e.dec[l];
    bf16* cache = dw.cache + (size_t)l * B * cache_bs;
    RmsFwdArgs r{};
    r.x = w.y[3 * l]; r.w = e.P + P.ln0; r.y_bf16 = w.dn1[l]; r.ld_bf16 = d; r.M = B; r.eps = c.eps; r.scale = 1.f;
    VQ_TRY(rmsnorm_fwd(r, st));
    // q|k|v of position t -> cache[b, t, :]
    VQ_TRY(gemm_fwd(w.dn1[l], d, e.W + P.qkv, d, cache + (size_t)t * 3 * d, (int)cache_bs, B, 3 * d, EPI_BF16, st));
    AttnArgs a{};
    a.q = cache + (size_t)t * 3 * d; a.k = cache + d; a.v = cache + 2 * d; a.ldq = a.ldk = a.ldv = 3 * d;
    a.q_bstride = a.k_bstride = a.v_bstride = cache_bs; a.q_off = t;
    a.o = w.dao[l]; a.ldo = d; a.lse = nullptr; a.B = B; a.H = H; a.Sq = 1; a.Sk = t + 1;
    a.rel_table = e.P + e.o_dec_rel; a.rel_bucket = e.dec_bucket; a.rel_mode = 2; a.causal = 1;
    VQ_TRY(attn_fwd(a, st));
    VQ_TRY(gemm_fwd(w.dao[l], d, e.W + P.o, d, w.y[3 * l + 1], d, B, d, EPI_RESID_F32, st, w.y[3 * l], d));
    r.x = w.y[3 * l + 1]; r.w = e.P + P.ln1; r.y_bf16 = w.dn2[l];
    VQ_TRY(rmsnorm_fwd(r, st));
    VQ_TRY(gemm_fwd(w.dn2[l], d, e.W + P.cq, d, w.cq[l], d, B, d, EPI_BF16, st));
    AttnArgs x{};
    x.q = w.cq[l]; x.ldq = d; x.k = w.kv_all + (size_t)l * 2 * d; x.v = x.k + d; x.ldk = x.ldv = ldkv;
    x.o = w.cao[l]; x.ldo = d; x.lse = nullptr; x.B = B; x.H = H; x.Sq = 1; x.Sk = S2;
    x.rel_mode = 0; x.keymask = w.cross_mask; x.causal = 0;
    VQ_TRY(attn_fwd(x, st));
    VQ_TRY(gemm_fwd(w.cao[l], d, e.W + P.co, d, w.y[3 * l + 2], d, B, d, EPI_RESID_F32, st, w.y[3 * l + 1], d));
    r.x = w.y[3 * l + 2]; r.w = e.P + P.ln2; r.y_bf16 = w.dn3[l];
    VQ_TRY(rmsnorm_fwd(r, st));
    VQ_TRY(gemm_fwd(w.dn3[l], d, e.W + P.wi, d, w.dh[l], f, B, f, EPI_RELU_BF16, st));
    VQ_TRY(gemm_fwd(w.dh[l], f, e.W + P.wo, f, w.y[3 * l + 3], d, B, d, EPI_RESID_F32, st, w.y[3 * l + 2], d));
  }
  RmsFwdArgs r{};
  r.x = w.y[3 * Ld]; r.w = e.P + e.o_dec_final; r.y_bf16 = w.yfin; r.ld_bf16 = d; r.M = B; r.eps = c.eps;
  r.scale = 1.f / sqrtf((float)d);
  VQ_TRY(rmsnorm_fwd(r, st));
  // tied LM head with the greedy argmax fused into the epilogue (fp32 accumulators -> per-tile maxima)
  VQ_TRY(gemm_fwd(w.yfin, d, e.W + e.o_shared, d, dw.pval, dw.slots, B, c.vocab_size, EPI_ARGMAX, st, dw.pidx, dw.slots));
  return 0;
}

}  // namespace vq

using namespace vq;

extern "C" int64_t vqacl_generate_workspace_bytes(void* engine, int B, int L, int N, int max_len) {
  (void)L; (void)N;
  return decode_carve(*reinterpret_cast<Engine*>(engine), nullptr, B, max_len, nullptr);
}

extern "C" int vqacl_generate(void* engine, const vqacl_batch* batch, const vqacl_proto_state* proto, int max_len,
                              int64_t* out_tokens, void* workspace, int64_t workspace_bytes, int* out_len, void* stream) {
  Engine& e = *reinterpret_cast<Engine*>(engine);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VQ_CHECK(batch && proto && out_tokens && out_len, "generate: null argument");
  VQ_CHECK(max_len >= 2 && max_len <= 64, "generate: max_len=%d must be in [2,64]", max_len);
  // the step workspace must be bound for (B, L, N, T >= 1); decode uses its decoder buffers with M = B rows
  VQ_CHECK(e.ws_base && batch->B == e.B && batch->L == e.L && batch->N == e.N && e.T >= 1,
           "generate: bind the step workspace for this (B, L, N) first");
  if (check_batch(e, batch, false)) return 1;
  DecodeWs dw;
  const int64_t need = decode_carve(e, reinterpret_cast<uint8_t*>(workspace), batch->B, max_len, &dw);
  VQ_CHECK(workspace && ((uintptr_t)workspace & 255) == 0 && workspace_bytes >= need,
           "generate: workspace of %lld bytes (256-byte aligned) required, %lld given", (long long)need, (long long)workspace_bytes);
  const vqacl_config& c = e.cfg;
  const int B = batch->B, S2 = batch->L + batch->N + 2, d = c.d_model, Ld = c.n_dec_layers;
  e.seed = 0;
  e.training = false;
  e.fwd_valid = false;
  VQ_TRY(encoder_forward(e, batch, st));
  vqacl_proto_state ps = *proto;
  ps.proto_update = 0;                                     // modeling_t5_our.py:607-612: frozen banks at test time
  VQ_TRY(si_path(e, batch, &ps, false, st));
  // encoder_forward waited for optimizer chunks 0..Le only; the cross-KV projection and every decode step read the
  // decoder weights, which a pending overlapped optimizer step writes last
  VQ_TRY(wait_params(e, 1 + c.n_enc_layers, st));
  VQ_TRY(gemm_fwd(e.w.mem, d, e.W + e.o_ckv, d, e.w.kv_all, Ld * 2 * d, B * S2, Ld * 2 * d, EPI_BF16, st));
  (void)vq_launch(decode_init_kernel, dim3((B + 255) / 256), dim3(256), 0, st, out_tokens, max_len, dw.cur, dw.unfinished, B, c.start_id);
  VQ_LAUNCH_CHECK();
  // rows-still-unfinished counter per generated column; the host looks at it every CHECK_EVERY tokens (one stream
  // synchronisation each) instead of after every token: finished rows emit pad, so running a few columns past the point
  // where HF's loop stops changes nothing in the columns that are returned
  constexpr int CHECK_EVERY = 4;
  static int* h_cnt = nullptr;
  if (!h_cnt) VQ_CUDA(cudaMallocHost(&h_cnt, 65 * sizeof(int)));
  VQ_CUDA(cudaMemsetAsync(dw.n_unfinished, 0, (size_t)(max_len + 1) * sizeof(int), st));
  const int nslots = ((c.vocab_size + 255) / 256) * 2;
  int len = 1, checked = 1;
  for (int t = 0; t + 1 < max_len; ++t) {
    VQ_TRY(decode_step(e, dw, B, S2, max_len, t, st));
    (void)vq_launch(decode_advance_kernel, dim3((B + 7) / 8), dim3(256), 0, st, (const float*)dw.pval, (const int*)dw.pidx, dw.slots, nslots,
                    out_tokens, max_len, t + 1, dw.cur, dw.unfinished, dw.n_unfinished, B, c.pad_id, c.eos_id);
    VQ_LAUNCH_CHECK();
    len = t + 2;
    if ((t + 1) % CHECK_EVERY == 0 || t + 2 == max_len) {
      VQ_CUDA(cudaMemcpyAsync(h_cnt, dw.n_unfinished, (size_t)(max_len + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
      VQ_CUDA(cudaStreamSynchronize(st));
      bool done = false;
      for (int col = checked; col <= t + 1; ++col)
        if (h_cnt[col] == 0) { len = col + 1; done = true; break; }   // every row had produced EOS after column `col`
      checked = t + 2;
      if (done) break;
    }
  }
  *out_len = len;
  return 0;
}
