// The VL-T5 / VQACL step engine: owns the parameter-arena layout and the activation workspace carve-up and issues
// every kernel of forward, backward, optimizer and greedy decode from C++ (no Python between launches).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../../include/vqacl_b200.h"
#include "gemm.h"
#include "ops.h"

namespace vq {

typedef __nv_bfloat16 bf16;

struct ParamInfo {
  std::string name;
  size_t off;
  int rows, cols;
  int group;  // 0 = weight-decay group, 1 = no-decay group ("bias" in the name), 2 = never receives a gradient
};

struct EncLayer { size_t ln0, qkv, o, ln1, wi, wo; };
struct DecLayer { size_t ln0, qkv, o, ln1, cq, co, ln2, wi, wo; };

struct Workspace {
  // encoder (M = B*S rows)
  bf16* feats_bf16; float* featpre;
  std::vector<float*> x;          // residual stream snapshots, 2*Le + 1
  std::vector<bf16*> n1, qkv, ao, n2, h;
  std::vector<uint32_t*> hmask, dhmask;                    // ReLU sign bitmasks [rows, ceil(d_ff/32)] written by the wi epilogue
  std::vector<float*> lse_e;
  float* enc_hidden;              // [B,S,d] fp32 (post final norm/dropout) -> encoder_hidden_states
  bf16* mem;                      // [B,S+2,d] decoder memory
  float *enc_mask, *cross_mask, *mask01;
  // SI path
  float *meanQ, *meanV, *curQ, *curV, *cntQ, *cntV, *pnorm;
  float *memdQ, *memdV, *mem_ssq, *mem_loss;   // memory_loss: mean - assigned prototype per sample, scratch, the two losses
  int64_t *idxQ, *idxV;
  // decoder (Md = B*T rows)
  int64_t* dec_ids;
  std::vector<float*> y;          // 3*Ld + 1
  std::vector<bf16*> dn1, dqkv, dao, dn2, cq, cao, dn3, dh;
  std::vector<float*> lse_s, lse_c;
  bf16* kv_all;                   // [B*(S+2), Ld*2*d]
  bf16* yfin;                     // [Md, d]
  bf16* logits;                   // [Md, ldv]
  float *lse_ce, *loss_rows, *w_rows, *loss;
  // backward scratch
  float *gd, *ge;                 // fp32 residual-stream gradients (decoder / encoder)
  // dY operands that the weight-gradient GEMMs read on the side stream live in small rings (3 layers deep) so that the
  // main stream can run ahead without write-after-read hazards (see backward())
  static constexpr int RING = 3;
  bf16* gdb_ring[3 * RING]; bf16* geb_ring[2 * RING];     // bf16 residual-stream gradients (masked for the consuming branch)
  bf16 *t_dqkv[RING], *t_dh[RING], *t_dcq[RING];          // decoder dY temporaries
  bf16 *t_eqkv[RING], *t_eh[RING];                        // encoder dY temporaries
  bf16 *t_d768, *t_e768;                                  // main-stream-only temporaries
  float* t_parts;                                         // [3][Md, d] fp32 slabs of the decoder FFN-out split-K GEMM
  float* t_d768_f32;                                      // split-K target of the LM-head dX GEMM
  bf16* dkv_all; bf16* dmem; bf16* dfeatpre;
  float* vis_partials;            // [num_sms, 10*768] per-CTA column sums of the visual-embedding backward
};

struct Engine {
  vqacl_config cfg;
  std::vector<ParamInfo> params;
  size_t n_decay = 0, n_train = 0, n_total = 0;
  std::vector<EncLayer> enc;
  std::vector<DecLayer> dec;
  size_t o_dec_final = 0, o_ckv = 0, o_enc_final = 0, o_Wf = 0, o_wf = 0, o_Wp = 0, o_wp = 0, o_img = 0, o_shared = 0;
  size_t o_bf = 0, o_bp = 0, o_enc_rel = 0, o_dec_rel = 0, o_tail = 0;
  float* P = nullptr; float* G = nullptr; bf16* W = nullptr;
  int enc_bucket_h[128] = {0}, dec_bucket_h[128] = {0};   // host copies of the rel -> bucket maps (go into kernel params)
  const int* enc_bucket = nullptr; const int* dec_bucket = nullptr;
  // workspace
  uint8_t* ws_base = nullptr; int64_t ws_bytes = 0;
  int B = 0, L = 0, N = 0, T = 0;
  Workspace w;
  std::map<std::string, int64_t> ws_names;
  int ldv = 0;  // logits pitch
  // step state
  uint32_t seed = 0; bool training = false; bool fwd_valid = false;
  bool mem_loss_valid = false;          // this forward computed memory_loss (its diffs are in the workspace)
  const float* mem_loss_g = nullptr;    // device float[2]: d(total loss) / d(loss_memory_Q, loss_memory_V) for the coming backward
  // side stream for the weight-gradient GEMMs (nobody consumes dW before the optimizer, so they run beside the dX chain)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_layer[4] = {nullptr, nullptr, nullptr, nullptr};
  int gdb_i = 0, geb_i = 0;
  // gradient arena zeroed ahead of backward on the side stream (during the decoder forward)
  bool g_prezeroed = false; bool prezero_request = false; cudaEvent_t ev_gzero = nullptr;
  // multi-GPU: per-stage completion events (main / side stream) that the host's communication stream waits on before the
  // stage callback issues the all-reduce of the gradient range that stage finalised (no main<->side join per stage)
  std::vector<cudaEvent_t> ev_stage_main, ev_stage_side;
  // optimizer overlapped with the next forward: chunks of the arena are updated on `opt_stream` in the order the forward
  // uses them (embeddings+visual | encoder layer 0..Le-1 | decoder+cross-KV); ev_opt[k] marks chunk k ready
  cudaStream_t opt_stream = nullptr;
  std::vector<cudaEvent_t> ev_opt;
  cudaEvent_t ev_opt_fork = nullptr;
  bool opt_pending = false;
  std::vector<cudaEvent_t> ext_ev;   // host-owned "chunk k gathered" events of the sharded optimizer (vqacl_set_param_events)
  bool ext_pending = false;
  // engine-owned scratch that must survive workspace re-binds (an overlapped optimizer step reads the gradient norm after
  // step() returned, possibly after the next train_step re-carved the workspace): [0,4) sumsq | [4,8) device error flags
  // | [64, 64 + OPT_PARTIALS) sum-of-squares partials
  static constexpr int OPT_PARTIALS = 16384;
  float* opt_scratch = nullptr;
  float* sumsq() const { return opt_scratch; }
  int* err_flags() const { return reinterpret_cast<int*>(opt_scratch + 4); }
  float* sumsq_partials() const { return opt_scratch + 64; }
  // decode workspace (separate carve, see decode.cu)
  Dropout drop(uint32_t site) const;
  int S() const { return L + N; }
};

#define VQ_TRY(expr) do { if ((expr) != 0) return 1; } while (0)

// Y[M,N] = epilogue(X[M,K] * W[N,K]^T): both operands K-major (nn.Linear forward)
inline int gemm_fwd(const bf16* A, int lda, const bf16* Wt, int K, void* C, int ldc, int M, int N, int epi, cudaStream_t st,
                    const void* R = nullptr, int ldr = 0, Dropout dr = Dropout(), float alpha = 1.f) {
  GemmArgs g{};
  g.epi = epi; g.M = M; g.N = N; g.K = K; g.C = C; g.ldc = ldc; g.R = R; g.ldr = ldr; g.alpha = alpha; g.splits = 1;
  g.drop_thr = dr.thr; g.drop_inv_keep = dr.inv_keep; g.seed = dr.seed; g.site = dr.site;
  return gemm_bf16(GemmOperand{A, lda, false}, GemmOperand{Wt, K, false}, g, 0, st);
}

int check_batch(const Engine& e, const vqacl_batch* b, bool need_labels);
// make `st` wait until optimizer chunk k (see Engine::ev_opt) has been written; no-op when no overlapped step is pending
int wait_params(Engine& e, int chunk, cudaStream_t st);
int encoder_forward(Engine& e, const vqacl_batch* b, cudaStream_t st);
int si_path(Engine& e, const vqacl_batch* b, const vqacl_proto_state* ps, bool sums_ready, cudaStream_t st);

int engine_build_layout(Engine& e);
int64_t engine_carve(Engine& e, uint8_t* base, int B, int L, int N, int T);  // base == nullptr: size query only

}  // namespace vq
