// kernels.h — launchers for the HBM-bound fused kernels and the attention kernels of the MedViLL pre-training step.
// `f32` selects the activation dtype: 0 = bf16 (production), 1 = fp32 (check mode). All launch on `stream`, return 0
// or a negative status (message via mv::last_error()).
#pragma once
#include "common.cuh"
#include "mask.cuh"

namespace mv {

// ---- joint embedding: [CLS] + regions + [SEP] + text  (models/cxrbert_origin.py:112-125, 22-35; upstream BertEmbeddings)
struct EmbedArgs {
  int B, L, H, N, T, A;              // A = N + 2, L = A + T
  const int64_t* cls_tok;          // [B]      int64 (reference dtype)
  const int64_t* sep_tok;          // [B]
  const int64_t* input_ids;        // [B, T]
  const int64_t* segment;          // [B, T]
  const int64_t* region_idx;       // [N]      sampled grid positions (models/image.py:64-69)
  const float* word; const float* pos; const float* type;   // fp32 master tables
  const float* gamma; const float* beta; float eps;
  const void* proj;                  // [B*N, H] activation dtype: img projection + bias (GEMM output)
  void* emb_sum;                     // [B*L, H] pre-LayerNorm sum (saved for backward)
  void* out;                         // [B*L, H] LN + dropout output = encoder input
  float* out32 = nullptr;            // optional fp32 copy of `out` (residual operand of layer 0 when the residual stream is fp32)
  int drop_on; uint32_t drop_site; DropoutCfg drop;
  DropoutCfg drop_img;               // region rows 1..N: ImageBertEmbeddings.dropout (args.dropout_prob); the rest: BertEmbeddings.dropout
  int V = 0, P = 0, TV = 0;          // table sizes (vocab, max_pos, type vocab) for the range checks; 0 = unchecked
  int* err = nullptr;                // device error-flag word (MV_ERR_*): out-of-range ids are clamped to 0 and reported
  int sep_pos = 0;                   // position id of the prefix [SEP]: 0 (pre-training, cxrbert_origin.py:119) or A-1
                                     // (fine-tune model, .../pytorch_pretrained_bert/model.py:886-892)
  int prefix_type = 0;               // token type of [CLS] / regions / [SEP]: 0, or 4 with new_segment_ids (data_loader.py:344)
};
int embed_ln_fwd(const EmbedArgs& a, int f32, cudaStream_t s);

struct EmbedBwdArgs {
  int B, L, H, N, T, A, V;
  const int64_t* cls_tok; const int64_t* sep_tok; const int64_t* input_ids; const int64_t* segment;
  const int64_t* region_idx;
  const void* dsum;                  // [B*L, H] gradient w.r.t. the pre-LN sum (output of ln_bwd)
  float* d_word; float* d_pos; float* d_type;   // fp32 gradient tables (atomically accumulated)
  void* d_proj;                      // [B*N, H] gradient of the projected image rows (activation dtype)
  int pad_id;                        // upstream nn.Embedding(padding_idx=0): lookup gradient of [PAD] is dropped
                                     // (-1: keep it — the vendored fine-tune BertEmbeddings has no padding_idx, model.py:228)
  int sep_pos = 0, prefix_type = 0;  // as EmbedArgs
  int TV = 2;                        // token-type vocabulary size (rows of d_type), <= 8
  int P = 0;                         // max_pos (0 = unchecked): ids outside the tables are clamped to 0 as in the forward
};
int embed_bwd_scatter(const EmbedBwdArgs& a, int f32, cudaStream_t s);

// ---- LayerNorm (eps is an argument: 1e-12 encoder, 1e-5 TF-style MLM head, models/cxrbert_origin.py:189-202)
// x_f32 (bf16 mode only): x is an fp32 pre-LN sum (fp32 residual stream); y32 (optional): fp32 copy of the output
int ln_fwd(const void* x, void* y, const float* gamma, const float* beta, int rows, int H, float eps, int drop_on,
           uint32_t drop_site, const DropoutCfg& drop, int f32, cudaStream_t s, int x_f32 = 0, float* y32 = nullptr);
// rows r with (r % period) in [lo, hi) use `drop` instead of ln_bwd's main DropoutCfg for in_drop (the region rows of the
// joint embedding have their own dropout probability)
struct LnAltDrop { int period, lo, hi; DropoutCfg drop; };
// dy: grad w.r.t. LN output (with in_drop: grad w.r.t. dropout(LN(x)), mask re-generated from drop_site)
// dx: grad w.r.t. x; dx_drop (optional, out_drop): dx * dropout mask of `drop_site` (grad of the dense output that
// was dropped before the residual add). dgamma/dbeta/dbias are fp32 and accumulated atomically (dbias <- dx_drop or dx).
int ln_bwd(const void* dy, const void* x, const float* gamma, void* dx, void* dx_drop, float* dgamma, float* dbeta,
           float* dbias, int rows, int H, float eps, int in_drop, int out_drop, uint32_t drop_site,
           const DropoutCfg& drop, int f32, cudaStream_t s, int x_f32 = 0, const struct LnAltDrop* alt = nullptr);

// ---- small utilities
int colsum_add(const void* x, long ld, int rows, int cols, float* out, int f32, cudaStream_t s);   // out[c] += sum_r x[r,c]
// dst[i,:] = src[(i / period) * stride + idx[i % period], :]
int gather_rows(const void* src, void* dst, const int64_t* idx, int n, int period, long stride, int H, int f32,
                cudaStream_t s);
// dst[(i / period) * stride + idx[i % period], :] (+)= src[i,:]
int scatter_rows(const void* src, void* dst, const int64_t* idx, int n, int period, long stride, int H, int add,
                 int f32, cudaStream_t s);
int dgelu_mul(const void* dy, const void* pre, void* dx, long n, int f32, cudaStream_t s);          // dx = dy * gelu'(pre)
int mul_elem(const void* a, const void* b, void* out, long n, int f32, cudaStream_t s);             // out = a * b
// d_pre (+)= d_out * (1 - out^2): backward of out = tanh(pre) (BertPooler)
int tanh_bwd(const void* d_out, const void* out, void* d_pre, long n, int add, int f32, cudaStream_t s);
int cast_f32_to_bf16(const float* src, bf16* dst, long n, cudaStream_t s);
int cast_bf16_to_f32(const bf16* src, float* dst, long n, cudaStream_t s);
int mask_dump(const unsigned char* mode, const int* t_len, int B, int A, int L, unsigned char* out, cudaStream_t s);
// derive (mode, t_len) from an explicit [B,L,L] (or [B,L]) int64 mask and count cells that differ from the predicate
int mask_classify(const int64_t* mask, int dims, int B, int A, int L, unsigned char* mode, int* t_len,
                  int* mismatches, cudaStream_t s);

// ---- losses (models/train_origin.py:62-63,118-126) and step metrics (:133-146)
struct CeArgs {
  int n, V; long ldv;
  const float* logits;               // [n, ldv] fp32
  const int64_t* labels;           // [n]
  void* dlogits;                     // [n, ldv] activation dtype: (softmax - onehot) * gscale, pad columns zeroed
  float gscale;                      // 1 / (#labelled tokens in the GLOBAL batch)
  const float* count_dev = nullptr;  // optional device scalar holding that count (all-reduced in-stream): gscale = 1 / *count_dev
  float* loss_sum;                   // += sum_i (lse_i - logit_i[label_i])
  int* correct;                      // += #(argmax == label)
  float* row_lse; int* row_argmax;   // optional per-row outputs (parity aids)
  const float* row_weight = nullptr; // optional [n] per-row loss weights (fine-tune masked_weights / multiplicities)
  float* row_loss = nullptr;         // optional [n] out: unweighted lse_i - logit_i[label_i]
  int* err = nullptr;                // device error-flag word: a label outside [0, V) is clamped to 0 and reported (MV_ERR_MLM_LABEL)
};
int mlm_ce_fwd_bwd(const CeArgs& a, int f32, cudaStream_t s);

// Luo's drop-worst (fine-tune model.py:1003-1010): per-sample loss = sum of weight * row_loss over the sample's rows; the
// `keep` samples with the smallest loss stay (ties: lower sample index first).  row_scale[i] = weight_i / (sum of kept
// weights + 1e-5) for rows of kept samples, 0 otherwise; *loss_sum += sum of kept losses / that denominator.
struct DropWorstArgs {
  int n, B, L, keep;
  const float* row_loss;             // [n]
  const float* row_weight;           // [n] or nullptr (= 1)
  const int64_t* rows;               // [n] flattened b * L + s
  float* row_scale;                  // [n] out
  float* loss_sum;
};
int drop_worst_select(const DropWorstArgs& a, cudaStream_t s);

struct ItmArgs {
  int B, H;
  const void* pooled;                // [B, H] activation dtype, tanh output
  const float* w; const float* b;    // [2, H], [2]
  const int64_t* labels;           // [B]
  float gscale;                      // 1 / (GLOBAL batch)
  const float* count_dev = nullptr;  // optional device scalar holding the global batch size: gscale = 1 / *count_dev
  float* logits;                     // [B, 2] fp32 out
  float* loss_sum; int* correct;
  void* d_pre;                       // [B, H] activation dtype: grad w.r.t. pooler pre-activation; null = forward only
  float* dw; float* db;              // fp32 grads (atomic)
  int no_tanh = 0;                     // 1: d_pre receives the gradient w.r.t. `pooled` itself (stand-alone ITM head), not w.r.t. the tanh input
  const float* ext_dlogits = nullptr;  // optional [B, 2]: gradient w.r.t. the logits supplied by the caller (autograd path) instead of the CE
};
int itm_head_fwd_bwd(const ItmArgs& a, int f32, cudaStream_t s);
// out[b] = softmax(logits[b, :2])[1]: the image-report match probability used as the retrieval similarity
// (Downstream_task/Retrieval/full_dset_retrieval.py:499-509)
int itm_match_prob(const float* logits, float* out, int B, cudaStream_t s);

// ---- optimizer: HF-3.x AdamW, correct_bias=True (models/train_origin.py:60,129-131)
struct AdamArgs {
  long n;
  float* p; float* g; float* m; float* v; bf16* shadow;
  float lr, beta1, beta2, eps, weight_decay; int step;
  float grad_scale;                  // multiply g before use (1 unless gradients were pre-accumulated unscaled)
  int zero_grad;                     // write zeros back to g (fuses optimizer.zero_grad())
};
int adamw_step(const AdamArgs& a, cudaStream_t s);

// ---- fine-tune optimizer: BertAdam (.../sc/pytorch_pretrained_bert/optimization.py:112-182): per-tensor gradient-norm
// clipping, Adam moments without bias correction, decoupled weight decay on non-bias / non-LayerNorm tensors
struct AdamChunk { long off; int n; int tensor; int decay; };   // a slice of ONE parameter tensor inside the arena
struct BertAdamArgs {
  float* p; float* g; float* m; float* v; bf16* shadow;
  const AdamChunk* chunks; int n_chunks;       // device
  float* sumsq; int n_tensors;                 // device scratch: squared gradient norm per tensor
  float lr, beta1, beta2, eps, weight_decay, max_grad_norm;
};
int bert_adam_step(const BertAdamArgs& a, cudaStream_t s);

// ---- BatchNorm2d of the frozen ResNet trunk, channels-last [rows, C]: batch statistics + affine (+ residual) (+ ReLU)
int bn_num_parts(long rows, int C);
long bn_workspace_floats(long rows, int C);
int bn_forward(const void* x, const void* resid, void* y, long rows, int C, const float* gamma, const float* beta,
               float* running_mean, float* running_var, float momentum, float eps, int training, int relu, float* workspace,
               long ws_floats, int f32, cudaStream_t s);

// uint8 [B,3,H,W] -> normalised channels-last activation [B,H,W,3] (data/helper.py:20-27 ToTensor + Normalize)
int normalize_u8(const unsigned char* src, void* dst, long B, long hw, int cpad, const float mean[3], const float stdv[3],
                 int f32, cudaStream_t s);
// same transform, written as 2x2 space-to-depth blocks [B, H/2 + 3, W/2 + 3, 16] (zero border 2 front / 1 back, channel =
// c*4 + dy*2 + dx, 12..15 zero): the input of the stem convolution in its 4x4 / stride-1 form (models/image.py)
int normalize_u8_s2d(const unsigned char* src, void* dst, int B, int H, int W, const float mean[3], const float stdv[3], int f32,
                     cudaStream_t s);
// stem tail: BatchNorm (train/eval) + ReLU + MaxPool2d(3, 2, 1) in one apply pass, channels-last [B,H,W,C] -> [B,H/2,W/2,C]
int bn_relu_maxpool(const void* x, void* y, int B, int H, int W, int C, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float momentum, float eps, int training, float* workspace, long ws_floats, int f32,
                    cudaStream_t s);

// ---- fused masked attention (upstream BertSelfAttention; twin .../pytorch_pretrained_bert/model.py:301-320)
struct AttnArgs {
  int B, L, nh, A;                   // head dim fixed at 64; H = nh * 64
  const unsigned char* mode;         // [B] MaskMode
  const int* t_len;                  // [B]
  const void* qkv;                   // [B*L, 3H] activation dtype (Q | K | V along columns)
  void* ctx;                         // [B*L, H]
  float* lse;                        // [B, nh, L] natural-log row log-sum-exp of the scaled, masked scores
  // backward only
  const void* dctx;                  // [B*L, H]
  void* dqkv;                        // [B*L, 3H]
  float* dq_acc;                     // [B*L, H] fp32 scratch (tcgen05 path; zeroed by the launcher)
  float* dq_part = nullptr;          // deterministic mode (tcgen05 path): [ceil(L/128)][B][L][H] fp32 per-key-tile dQ partials, summed
                                     // in key-tile order by the convert pass instead of fp32 reduce-adds in arrival order
  float* delta;                      // [B, nh, L] scratch: rowsum(dO * O)
  int drop_on; uint32_t drop_site; DropoutCfg drop;
#ifdef MV_ATTN_TIMELINE
  unsigned long long* timeline;      // [4][64] clock64() marks of one probe CTA (tests/attn_timeline build only)
#endif
};
int attention_fwd_tc05(const AttnArgs& a, cudaStream_t s);
int attention_bwd_tc05(const AttnArgs& a, cudaStream_t s);
int attention_fwd_simt(const AttnArgs& a, cudaStream_t s);
int attention_bwd_simt(const AttnArgs& a, cudaStream_t s);

}  // namespace mv
