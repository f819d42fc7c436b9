// Host-callable launchers of every non-GEMM kernel on the VL-T5 hot path (internal C++ interface; the C-ABI in
// include/vqacl_b200.h wraps these). All pointers are device pointers; all launches are asynchronous on `stream`.
#pragma once
#include "common.cuh"

namespace vq {

int num_sms();  // gemm.cu

constexpr int DM = 768;  // d_model of t5-base (the kernels are specialised for it; checked at engine creation)

struct Dropout {
  uint32_t thr = 0;       // 16-bit keep threshold round(p * 65536) (vq_dropout_pair), 0 disables
  float inv_keep = 1.f;   // 1 / (1 - p)
  uint32_t seed = 0;      // per-launch key = mix(step seed, site id), computed on the host (Engine::drop)
  uint32_t site = 0;      // informational
};

// ---------------------------------------------------------------- attention (attention.cu)
struct AttnArgs {
  const __nv_bfloat16 *q, *k, *v;
  int ldq, ldk, ldv;        // row pitches in elements; head h lives at columns [h*64, h*64+64)
  __nv_bfloat16* o;         // fwd out / bwd: unused
  int ldo;                  // pitch of o and dO
  float* lse;               // [B,H,Sq] fwd out / bwd in
  int B, H, Sq, Sk;
  const float* rel_table;   // [num_buckets, H] fp32 (relative_attention_bias.weight) or null
  const int* rel_bucket;    // HOST pointer, int32[127]: bucket of (key - query + 63), precomputed with the HF formula;
                            // copied into the kernel parameters (constant bank) so the bias table needs one global load per entry
  int rel_mode;             // 0 none, 1 bias on the text x text corner only (encoder), 2 bias everywhere (decoder self)
  int Lt;                   // text length for rel_mode 1
  const float* keymask;     // [B,Sk] additive key mask or null
  int causal;               // add -10000 where key > query (HF 4.2.1 decoder extended mask)
  uint32_t drop_thr; float drop_inv_keep; uint32_t seed, site;   // dropout on the probabilities
  // backward only
  const __nv_bfloat16* dO;
  const __nv_bfloat16* o_saved; // forward output (pitch ldo); when given and Sq <= 16 < Sk the key-split backward is used
  __nv_bfloat16 *dq, *dk, *dv;
  int lddq, lddk, lddv;
  float* d_rel_table;       // [num_buckets, H] accumulated with atomics, or null
  // KV-cached decoding (forward only): element strides between batch entries (0 = S * ld, i.e. densely packed) and the
  // absolute position of query row 0 (the relative-position bias of HF's position_bias[:, :, -seq_length:, :])
  long long q_bstride, k_bstride, v_bstride, o_bstride;
  int q_off;
};
int attn_fwd(const AttnArgs& a, cudaStream_t stream);
int attn_bwd(const AttnArgs& a, cudaStream_t stream);
// shared by attention.cu and attention_tc.cu
struct AttnBuckets { int8_t b[128]; };   // rel -> bucket map (rel = key - query + 63), passed by value (constant bank)
// dropout pair index of probabilities (q, k), (q, k+1) of problem `blk` = batch * H + head (k even), single-tile kernels
VQ_DEVINL uint32_t attn_pair_idx(uint32_t blk, int q, int k) { return ((blk * 64u + (uint32_t)q) * 64u + (uint32_t)k) >> 1; }
// attention_tc.cu: encoder self-attention on tcgen05 / TMEM / TMA
bool attn_tc_eligible(const AttnArgs& a);
int attn_enc_fwd_tc(const AttnArgs& a, const AttnBuckets& bk, cudaStream_t stream);
bool attn_tc_bwd_eligible(const AttnArgs& a);
int attn_enc_bwd_tc(const AttnArgs& a, const AttnBuckets& bk, cudaStream_t stream);

// ---------------------------------------------------------------- elementwise.cu
int cast_f32_to_bf16(const float* src, __nv_bfloat16* dst, size_t n, cudaStream_t stream);

// T5 RMSNorm forward: y = x * rsqrt(mean(x^2) + eps) * w * scale, optional dropout on y.
// Row r of x maps to output row (r / in_rpb) * out_rpb + r % in_rpb when in_rpb > 0 (used to write the encoder
// output straight into the [B, S+2, d] decoder-memory buffer), identity otherwise.
struct RmsFwdArgs {
  const float* x; const float* w;
  __nv_bfloat16* y_bf16; int ld_bf16;   // may be null
  float* y_f32; int ld_f32;             // may be null
  int M; float eps; float scale;
  int in_rpb, out_rpb;
  Dropout drop;
  // optional: the input row is formed here, x = resid + dropout_r(sum over n_parts slabs of parts, fixed order), and also
  // written to x_out (fp32 [M, 768]) — the deterministic meeting point of a split-K GEMM whose result is a residual update
  const float* parts; int n_parts; long long part_stride;
  const float* resid; float* x_out; Dropout resid_drop;
};
int rmsnorm_fwd(const RmsFwdArgs& a, cudaStream_t stream);

// RMSNorm backward fused with the residual-gradient add:
//   dy = dn * scale (* own dropout mask);  dx = rstd * (dy*w - xhat * mean(dy*w*xhat));  dw += sum_rows dy * xhat
//   g_out = g_in + dx;   gb_out = bf16(g_out * consumer dropout mask)
struct RmsBwdArgs {
  const __nv_bfloat16* dn; int ld_dn;   // upstream gradient (bf16)
  const float* dn_f32;                  // ... or fp32 (same pitch ld_dn), used when non-null
  int dn_zero;                          // fp32 form only: write zeros back after reading — dn_f32 is a split-K accumulation target
                                        // (red.global.add) that the next GEMM expects cleared, and this kernel is its only reader
  const __nv_bfloat16* dn2; int ld_dn2; // optional second upstream gradient added to dn (bf16), same row map
  int in_rpb, out_rpb;                  // row map applied to dn/dn2 rows (see RmsFwdArgs), identity if in_rpb == 0
  const float* x; const float* w;
  const float* g_in;                    // may be null
  float* g_out;                         // may be null (then only dw / gb_out)
  __nv_bfloat16* gb_out;                // may be null
  float* dw;                            // [768], accumulated
  int M; float eps; float scale;
  // optional per-sample gradient broadcast over the tokens of each side of the Q/V split (memory_loss backward): row r of
  // batch element b = r / bc_S gets bc_g[0] * bc_cq * bc_q[b] when r % bc_S < bc_split, else bc_g[1] * bc_cv * bc_v[b]
  const float* bc_q; const float* bc_v; const float* bc_g; float bc_cq, bc_cv; int bc_S, bc_split;
  Dropout own;                          // dropout that was applied to this norm's output (final norms)
  Dropout consumer;                     // dropout of the residual branch that consumes gb_out
  int consumer_cols;                    // element index = row * consumer_cols + col  (== 768)
};
int rmsnorm_bwd(const RmsBwdArgs& a, cudaStream_t stream);

// token embedding gather (+ dropout) into rows [row0, row0+L) of each batch element of x [B, S, 768]
// ids outside [0, vocab) read row 0 and set bit 0 of *err (device int, may be null); the host raises at its next error check
int embed_fwd(const int64_t* ids, int B, int L, const float* table, float* x, int S, int row0, Dropout drop, int vocab, int* err,
              cudaStream_t stream);
// scatter-add of g rows into dtable (same mapping/dropout as embed_fwd)
int embed_bwd(const int64_t* ids, int B, int L, const float* g, int S, int row0, float* dtable, Dropout drop, int vocab, cudaStream_t stream);
// decoder_input_ids = shift_right(labels) (start 0, -100 -> pad 0)
int shift_right(const int64_t* labels, int64_t* dec_ids, int B, int T, int start_id, int pad_id, cudaStream_t stream);
// additive key masks from input_ids: enc [B,S] (-10000 on text pads), cross [B,S+2] (-1e9 on text pads)
int build_keymasks(const int64_t* ids, int B, int L, int S, int pad_id, float* enc_mask, float* cross_mask, float* mask01,
                   cudaStream_t stream);   // mask01 [B,S+2]: the 1/0 encoder_attention_mask the reference returns

// box normalisation + clamp and one-hot labels on the device (vqa_data_memory.py:179-187, 386-393)
int collate_device(const float* boxes_px, const float* wh, int B, int N, float* boxes_out, const int64_t* cate_ids, int n_cate,
                   float* cate_oh, const int64_t* ques_ids, int n_ques, float* ques_oh, cudaStream_t stream);

// VisualEmbedding (modeling_t5_our.py:93-143) after the 2048->768 GEMM: bias + RMSNorm, box/area projection + RMSNorm,
// image-order and object-order embeddings, dropout; writes rows [L, L+N) of x [B,S,768].
struct VisArgs {
  const float* featpre;       // [B*N, 768] = feats @ Wf^T (no bias)
  const float* boxes;         // [B, N, 4]
  const float *bf, *wf;       // feat_embedding.0.bias, feat_embedding.1.weight
  const float *Wp, *bp, *wp;  // absolute_vis_pos_embedding.0.{weight [768,5], bias}, .1.weight
  const float* img_emb;       // img_order_embedding.weight [n_images, 768] (row 0 used)
  const float* shared;        // shared.weight [V, 768] (rows V-1-n used)
  int V, B, N, S, L;
  float eps;
  float* x;                   // [B, S, 768]
  Dropout drop;
  // backward
  const float* g;             // [B, S, 768] gradient w.r.t. x
  __nv_bfloat16* dfeatpre;    // [B*N, 768] bf16 (A operand of the dWf GEMM)
  float *dbf, *dwf, *dWp, *dbp, *dwp, *dimg, *dshared;
  float* partials;            // scratch [num_sms, 10*768] fp32 (per-CTA column sums, reduced in a fixed order)
};
int vis_embed_fwd(const VisArgs& a, cudaStream_t stream);
int vis_embed_bwd(const VisArgs& a, cudaStream_t stream);

// ---------------------------------------------------------------- proto.cu  (SI prototype path)
// token means of the encoder output: meanQ[b] = mean(h[b, :split]), meanV[b] = mean(h[b, split:S])
int proto_means(const float* h, int B, int S, int split, float* meanQ, float* meanV, cudaStream_t stream);
// calculate_current_prototype: proto[c] = sum_b labels[b,c] * mean[b] / (cnt[c] <= 0 ? 1 : cnt[c]); cnt[c] = sum_b labels[b,c]
int proto_scatter_mean(const float* mean, const float* labels, int B, int C, float* proto, float* cnt, cudaStream_t stream);
// finish a scatter-mean whose sums/counts were all-reduced across ranks: proto = sums / max-style divisor
int proto_div(float* proto_sums, const float* cnt, int C, cudaStream_t stream);
// same as proto_scatter_mean but leaves proto as un-divided sums (for the cross-rank all-reduce)
int proto_scatter_sum(const float* mean, const float* labels, int B, int C, float* proto, float* cnt, cudaStream_t stream);
// update_prototype state machine on explicit buffers (see kernel for the mode table)
struct ProtoUpdateArgs {
  const float *curQ, *curV, *cntQ, *cntV;
  float *Qproto, *Vproto, *numQ, *numV;
  int CQ, CV;
  int task_id;
  int first_step_of_task;   // current_task_id not in Q_task_cur_proto
  int has_mem;              // current_task_id in Q_task_mem_proto
  float alpha, beta;
};
int proto_update(const ProtoUpdateArgs& a, cudaStream_t stream);
// cosine_similarity_multi + feature mix: idx[b] = first argmax_c cos(tanh P_c, tanh x_b); out row = raw P[idx[b]] (bf16)
// scratch: [C,768] fp32 (normalised tanh of the bank)
int proto_retrieve(const float* P, int C, const float* x, int B, __nv_bfloat16* out, int out_pitch_rows, int out_row,
                   int64_t* idx, float* out_f32, float* scratch, cudaStream_t stream);

// memory_loss (nextqa/modeling_t5_nextqa.py:544-555): loss2[0] = mean_b |meanQ[b] - ques_labels[b] @ PQ|^2, loss2[1] the same for
// the V side; diffQ / diffV [B,768] (fp32) are kept for the backward pass, ssq is scratch [2*B]
int proto_memory_loss(const float* meanQ, const float* meanV, const float* ques_labels, const float* cate_labels, const float* PQ,
                      const float* PV, int CQ, int CV, int B, float* diffQ, float* diffV, float* ssq, float* loss2, cudaStream_t stream);

// ---------------------------------------------------------------- lmhead_ce.cu
// per-row log-sum-exp and CE loss over bf16 logits [M, V] (pitch ld); label -100 -> loss 0
int ce_fwd(const __nv_bfloat16* logits, int ld, int M, int V, const int64_t* labels, float* lse, float* loss, cudaStream_t stream);
// in place: logits <- (softmax - onehot) * w[row]   (w = dL/dloss_row; 0 for ignored rows)
int ce_bwd(__nv_bfloat16* logits, int ld, int M, int V, const int64_t* labels, const float* lse, const float* w, const float* gscale,
           cudaStream_t stream);   // gscale: optional device scalar multiplied into every w[row]
// fused loss tail of VLT5VQA.train_step (vqa_model.py:46-54): loss = mean_b(score_b * sum_t loss_bt / max(n_b,1));
// also emits the per-row weights w[b,t] = score_b / (max(n_b,1) * B) for valid labels
int loss_tail(const float* loss_rows, const int64_t* labels, const float* scores, int B, int T, float* loss_out, float* w_rows, cudaStream_t stream);
// greedy argmax over bf16 logits rows (first max), with the finished-row bookkeeping of HF greedy_search
int argmax_rows(const __nv_bfloat16* logits, int ld, int M, int V, int64_t* out, cudaStream_t stream);

// ---------------------------------------------------------------- optim.cu
// sum of squares of g[0:n) -> *out (fp32), two-stage deterministic
int grad_sumsq(const float* g, size_t n, float* partials, float* out, cudaStream_t stream);
int grad_sumsq_ranges(const float* g, const int64_t* begin, const int64_t* end, int n_ranges, float* partials, int partials_cap,
                      float* out, cudaStream_t stream);
// HF-4.2.1 AdamW over a flat arena: elements [0, n_decay) get weight decay, [n_decay, n) do not.
// clip coefficient = min(1, max_norm / (sqrt(*sumsq) + 1e-6)) when sumsq != null and max_norm > 0.
struct AdamArgs {
  float* p; const float* g; float* m; float* v; __nv_bfloat16* p_bf16;
  size_t n, n_decay;
  float lr, beta1, beta2, eps, weight_decay;
  int step;                 // 1-based
  const float* sumsq; float max_norm;
  int max_blocks;           // 0 = fill the GPU; > 0 caps the grid (overlapped mode: leave the SMs' thread slots to the forward's GEMMs)
};
int adamw_hf(const AdamArgs& a, cudaStream_t stream);

}  // namespace vq
