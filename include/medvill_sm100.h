/* medvill_sm100.h — C ABI of libmedvill_sm100.so: the B200-native (sm_100a) implementation of the MedViLL joint
 * vision-language PRE-TRAINING STEP.  Plain pointers and sizes only; no torch types.  Every entry returns 0 on success
 * or a negative status (message: mv_last_error()); nothing throws or exits.  All device pointers are caller-owned
 * (PyTorch-allocated) and only borrowed while the enqueued work runs; the handle owns workspaces, TMA descriptors,
 * events and the NCCL communicator.  Work is enqueued on the cudaStream_t passed as `stream` (void* here so the
 * header needs no CUDA include).  There is no CPU fallback: without an sm_100 GPU every compute entry fails loudly.
 *
 * Reference interfaces replaced (paths relative to the reference repo root):
 *   mv_forward / mv_backward        CXRBERT.forward + loss.backward()   models/cxrbert_origin.py:144-149,
 *                                   CXRBertEncoder.forward :87-130, ImageBertEmbeddings.forward :22-35,
 *                                   BertPreTrainingHeads :205-248, ImageTextMatching :164-173,
 *                                   losses models/train_origin.py:62-63,118-126, metrics :133-146
 *   mv_adamw_step                   transformers AdamW.step (upstream)  models/train_origin.py:60,129-131
 *   mv_bert_adam_step               BertAdam.step (fine-tune)           Downstream_task/report_generation_and_vqa/sc/
 *                                                                       pytorch_pretrained_bert/optimization.py:112-182
 *   mv_comm_* / bucketed all-reduce nn.DataParallel gradient reduction  models/train_origin.py:53-55
 *   mv_attn_mask_dump / classify    CXRDataset mask construction        data/dataset_origin.py:138-176
 *                                   get_extended_attn_mask              models/cxrbert_origin.py:75-85
 *   mv_itm_match_prob               retrieval similarity                Downstream_task/Retrieval/retrieval.py:29-32,
 *                                                                       full_dset_retrieval.py:499-509
 *   mv_full_logits                  prediction_scores [B,L,V] of CXRBERT.forward (drop-in output)
 *   mv_gemm / mv_attention_* / mv_layernorm_* / mv_mlm_ce             per-op entry points used by the parity tests
 */
#ifndef MEDVILL_SM100_H_
#define MEDVILL_SM100_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MV_ABI_VERSION 5   /* 2: mv_batch grew the embedding-layout switches, global_counts and lab_weights; new entries
                              3: mv_batch.drop_worst_keep
                              4: mv_config.{attn_dropout_p, img_dropout_p, flags}; mv_gemm_desc.resid_f32;
                                 mv_step_stats.error_flags; mv_backward_external, mv_itm_head, mv_profile_read
                              5: mv_stem_conv_s2d */

enum { MV_PREC_BF16 = 0, MV_PREC_FP32 = 1 };          /* activation / GEMM-operand precision policy */
enum { MV_MODE_BIDIR = 0, MV_MODE_S2S = 1, MV_MODE_BAR = 2, MV_MODE_NONCROSS = 3,      /* attention-mask modes */
       MV_MODE_S2S_FT = 4, MV_MODE_BAR_FT = 5 };   /* fine-tune variants: causal block ends at the real text end, padded
                                                      rows see the prefix only (.../sc/data_loader.py:394-408)          */
enum {                                                 /* GEMM epilogues (mv_gemm) */
  MV_EPI_NONE = 0, MV_EPI_BIAS = 1, MV_EPI_BIAS_GELU = 2, MV_EPI_BIAS_RESID = 3, MV_EPI_BIAS_TANH = 4,
  MV_EPI_RESID = 5, MV_EPI_DGELU = 6,
  MV_EPI_BIAS_GELU_GRAD = 7,   /* C = gelu(acc + bias), C2 = gelu'(acc + bias): saves the derivative instead of the pre-activation */
  MV_EPI_MUL = 8               /* C = acc * aux[m,n] (the backward partner of MV_EPI_BIAS_GELU_GRAD)                              */
};

typedef struct mv_handle mv_handle;

typedef struct mv_config {
  int32_t hidden, heads, layers, inter, vocab, max_pos, type_vocab;  /* BertConfig (head dim must be 64)            */
  int32_t num_image_embeds;   /* N: sampled regions            main_origin.py:137                                  */
  int32_t seq_len;            /* S: text tokens                main_origin.py:130  (T = S+1, L = N + S + 3)        */
  int32_t img_hidden;         /* 2048                          main_origin.py:133                                  */
  int32_t grid;               /* ResNet regions per image: (img_size/32)^2                                         */
  int32_t max_batch;          /* per-micro-batch sample capacity (workspaces are sized for it)                     */
  int32_t precision;          /* MV_PREC_*                                                                         */
  float ln_eps;               /* encoder / embedding LayerNorm eps (1e-12)                                         */
  float head_ln_eps;          /* MLM-head TF-style LayerNorm eps (1e-5)  models/cxrbert_origin.py:212              */
  float dropout_p;            /* hidden-state dropout: BertConfig.hidden_dropout_prob (text / [CLS] / [SEP]        */
                              /* embeddings and both dense outputs of every layer; upstream BertEmbeddings,        */
                              /* BertSelfOutput, BertOutput)                                                       */
  float attn_dropout_p;       /* attention-probability dropout: BertConfig.attention_probs_dropout_prob            */
  float img_dropout_p;        /* image-embedding dropout: args.dropout_prob  models/cxrbert_origin.py:19,33        */
  int32_t flags;              /* MV_FLAG_*                                                                         */
} mv_config;

enum {
  MV_FLAG_BF16_RESIDUAL = 1,  /* bf16 mode only: keep the residual stream (pre-LayerNorm sums, LayerNorm outputs that */
                              /* feed the next residual add) in bf16 instead of fp32 (A/B measurements; the default   */
                              /* fp32 stream is what meets the 1e-2 logit tolerance at 12 layers)                     */
  MV_FLAG_DETERMINISTIC = 2   /* attention backward: ordered dQ reduction (bit-reproducible; utils.set_seed turns it  */
                              /* on, as the reference sets cudnn.deterministic, utils/utils.py:9-16)                  */
};

/* Offsets (in elements) of every trainable tensor inside the flat parameter arena.  The fp32 master parameters,
 * their gradients, both Adam moments and the bf16 shadow all use this one layout, so the torch nn.Parameters of the
 * drop-in modules are views into the arena and gradient buckets are contiguous ranges.  Per-layer offsets are
 * relative to layer0 + l * layer_stride. */
typedef struct mv_layout {
  int64_t total;
  int64_t word, pos, type, emb_ln_g, emb_ln_b, img_w, img_b;
  int64_t layer0, layer_stride;
  int64_t l_wqkv, l_bqkv, l_wo, l_bo, l_ln1_g, l_ln1_b, l_w1, l_b1, l_w2, l_b2, l_ln2_g, l_ln2_b;
  int64_t pool_w, pool_b, mlm_bias, mlm_tw, mlm_tb, mlm_ln_g, mlm_ln_b, itm_w, itm_b;
  int64_t vocab_padded;       /* row stride of logits buffers (vocab rounded up to 64)                              */
} mv_layout;

typedef struct mv_batch {
  int32_t B;                   /* samples in this micro-batch (<= max_batch)                                        */
  const int64_t* cls_tok;      /* [B]      device                                                                   */
  const int64_t* sep_tok;      /* [B]                                                                               */
  const int64_t* input_ids;    /* [B, T]                                                                            */
  const int64_t* segment;      /* [B, T]                                                                            */
  const int64_t* is_aligned;   /* [B]      ITM labels                                                               */
  const int64_t* region_idx;   /* [N]      sorted sampled grid positions (models/image.py:64-69)                    */
  const uint8_t* mode;         /* [B]      MV_MODE_* per sample (the Mixed mode draws per sample)                   */
  const int32_t* t_len;        /* [B]      real text length incl. [SEP]                                             */
  const void* feats;           /* [B, grid, img_hidden] activation dtype: ResNet-50 grid features, channels-last    */
  int32_t n_lab;               /* number of labelled (MLM) positions in this micro-batch                            */
  const int64_t* lab_rows;     /* [n_lab]  flattened b * L + s of every position with txt_labels != -100            */
  const int64_t* lab_labels;   /* [n_lab]  the labels at those positions                                            */
  float inv_n_lab_global;      /* 1 / #labelled tokens of the GLOBAL batch (all ranks, all micro-batches)           */
  float inv_batch_global;      /* 1 / GLOBAL batch size                                                             */
  uint64_t dropout_seed;       /* per-step seed of the counter-based dropout RNG                                    */
  int32_t train;               /* 1: dropout active (if dropout_p > 0) and backward state is kept                   */
  /* embedding-layout switches; all 0 = the pre-training model (models/cxrbert_origin.py:112-125).  The report-generation
   * fine-tune model (Downstream_task/report_generation_and_vqa/sc/pytorch_pretrained_bert/model.py:864-900,223-260)
   * differs in exactly these three places: */
  int32_t sep_position;        /* position id of the prefix [SEP]: 0, or A-1 (= N+1) in the fine-tune model        */
  int32_t prefix_type;         /* token type of [CLS]/regions/[SEP]: 0, or 4 with new_segment_ids                   */
  int32_t pad_lookup_grad;     /* 1: keep the embedding-lookup gradient of [PAD] (vendored nn.Embedding: no padding_idx) */
  const float* global_counts;  /* optional DEVICE [2] = {#labelled tokens, batch size} of the GLOBAL batch, e.g. all-reduced     */
                               /* in-stream by mv_comm_allreduce_f32; when set it replaces inv_n_lab_global / inv_batch_global */
                               /* so that a multi-rank step needs no host round trip for the loss normalisers             */
  const float* lab_weights;    /* [n_lab] optional per-row loss weights (fine-tune masked_weights, model.py:998-1005;  */
                               /* a position masked twice is one row of weight 2); NULL = 1                            */
  int32_t drop_worst_keep;     /* > 0: Luo's drop-worst of the fine-tune loss (model.py:1003-1010): only the                */
                               /* drop_worst_keep = int(B * (1 - drop_worst_ratio)) samples with the smallest weighted loss   */
                               /* contribute (an integer, so the host rounds exactly as the reference's Python does);         */
                               /* mlm_loss_sum then receives the NORMALISED loss (kept sum / (kept weights + 1e-5)) because   */
                               /* the denominator depends on which samples were kept; global_counts[0] is not used and        */
                               /* inv_n_lab_global is a plain multiplier on the gradient (1, or 1 / world when ranks average   */
                               /* their gradients as DistributedDataParallel does).  0 = every sample counts                  */
} mv_batch;

typedef struct mv_step_stats {
  float mlm_loss_sum;          /* sum over labelled tokens of the token CE  (mean = sum * inv_n_lab)                */
  float itm_loss_sum;          /* sum over samples of the ITM CE                                                    */
  int32_t mlm_correct, itm_correct;   /* models/train_origin.py:133-146                                             */
  int32_t error_flags;         /* MV_ERR_*: set by the kernels when an input is out of range (the id is clamped, the step   */
                               /* continues); PyTorch raises IndexError in the same situations                            */
} mv_step_stats;
enum { MV_ERR_TOKEN_ID = 1, MV_ERR_SEGMENT_ID = 2, MV_ERR_REGION_IDX = 4, MV_ERR_MLM_LABEL = 8 };

/* Gradients computed by the CALLER for mv_backward_external: the torch.autograd drop-in path runs the reference's own loss
 * code on the outputs of CXRBERT.forward (models/train_origin.py:118-130) and hands d(loss)/d(outputs) back.  Every pointer is
 * optional (NULL = that output received no gradient). */
typedef struct mv_external_grads {
  int32_t n_rows;              /* rows of the [B*L, V] prediction scores whose gradient is not identically zero              */
  const int64_t* rows;         /* [n_rows] flattened b * L + s, ascending (device)                                            */
  const void* dlogits;         /* [n_rows, vocab_padded] activation dtype: gradient of those rows, pad columns zero (device)  */
  const float* d_itm;          /* [B, 2] fp32: gradient of the ITM logits (device)                                            */
  const void* d_seq;           /* [B*L, hidden] activation dtype: gradient of the final hidden states (CXRBertEncoder output) */
  const void* d_pooled;        /* [B, hidden] activation dtype: gradient of the pooled output (CXRBertEncoder output)         */
} mv_external_grads;

typedef struct mv_gemm_desc {
  int32_t M, N, K;
  const void* A; int64_t lda; int32_t a_mn;   /* a_mn = 0: A is [M,K] row-major; 1: stored [K,M]                    */
  const void* B; int64_t ldb; int32_t b_mn;   /* b_mn = 0: B is [N,K] row-major (nn.Linear weight); 1: stored [K,N] */
  void* C; int64_t ldc; int32_t c_f32; int32_t accumulate;
  void* C2; int64_t ldc2;
  int32_t epi;
  const float* bias;
  const void* resid; int64_t ldr;
  const void* aux; int64_t ldaux;
  float dropout_p; uint64_t dropout_seed; uint32_t dropout_site;
  int32_t resid_f32;           /* MV_EPI_BIAS_RESID / MV_EPI_RESID with an fp32 residual operand and fp32 C (c_f32 = 1)    */
} mv_gemm_desc;

const char* mv_last_error(void);
int mv_abi_version(void);

/* ---- lifecycle ---- */
int mv_layout_query(const mv_config* cfg, mv_layout* out);               /* host only, no GPU needed                */
int mv_bucket_plan(const mv_config* cfg, int64_t* offsets, int64_t* counts, int32_t max_buckets, int32_t* n_buckets);
int mv_create(mv_handle** out, const mv_config* cfg);
int mv_destroy(mv_handle* h);
int mv_bind_arenas(mv_handle* h, float* params, float* grads, float* adam_m, float* adam_v, void* shadow_bf16);
int mv_refresh_shadow(mv_handle* h, void* stream);                       /* bf16 shadow <- fp32 master              */

/* ---- the pre-training step ---- */
int mv_stats_reset(mv_handle* h, void* stream);
int mv_forward(mv_handle* h, const mv_batch* b, void* stream);           /* fwd + both losses + metrics             */
int mv_backward(mv_handle* h, const mv_batch* b, int32_t allreduce, void* stream);  /* grads += ; optional bucketed */
                                                                         /* NCCL all-reduce overlapped with bwd     */
/* backward from caller-supplied output gradients (grads +=), same bucketed all-reduce option as mv_backward */
int mv_backward_external(mv_handle* h, const mv_batch* b, const mv_external_grads* g, int32_t allreduce, void* stream);
int mv_zero_grads(mv_handle* h, void* stream);
int mv_adamw_step(mv_handle* h, float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                  float grad_scale, void* stream);                       /* waits for pending all-reduces; zeroes g */
/* BertAdam.step of the report-generation fine-tune (Downstream_task/report_generation_and_vqa/sc/pytorch_pretrained_bert/
 * optimization.py:112-182; parameter groups finetune.py:383-395): per-parameter clip_grad_norm_(p, max_grad_norm), Adam
 * moments without bias correction, update += weight_decay * p for non-bias / non-LayerNorm tensors, p -= lr * update with
 * `lr` already scheduled by the caller (warmup_linear, optimization.py:45-48).  Pooler and ITM head are skipped (no
 * gradient on that path).  Zeroes the gradients it consumed and refreshes the bf16 shadow. */
int mv_bert_adam_step(mv_handle* h, float lr, float beta1, float beta2, float eps, float weight_decay, float max_grad_norm,
                      void* stream);
int mv_read_stats(mv_handle* h, mv_step_stats* host_out, void* stream);  /* D2H + stream sync                       */
int mv_read_stats_async(mv_handle* h, mv_step_stats* pinned_out, void* stream);  /* D2H enqueue only (pinned dst):   */
                                                                         /* lets the trainer log the loss lazily    */
                                                                         /* (train_origin.py:118-146 syncs per step)*/
int mv_itm_logits(mv_handle* h, float* host_out, int32_t B, void* stream);
/* softmax(itm_logits)[:, 1] of the last mv_forward -> device_out[B] (device pointer, stream-ordered, no sync): the pair
 * similarity of label-conditioned retrieval, CXRBertForRetrieval.forward + nn.Softmax(dim=1)(logits)[:, 1]
 * (Downstream_task/Retrieval/retrieval.py:29-32, full_dset_retrieval.py:499-509). */
int mv_itm_match_prob(mv_handle* h, float* device_out, int32_t B, void* stream);
int mv_full_logits(mv_handle* h, const mv_batch* b, float* logits, int64_t ld, void* stream);  /* [B*L, ld] fp32    */
int mv_peek(mv_handle* h, const char* name, int32_t layer, void* dst, int64_t max_bytes, int64_t* bytes, void* stream);

/* ---- measurement aids (bench.py): launch counter; CUDA-event timing per kernel family on the launch stream ---- */
long mv_launch_count(void);                                              /* kernels launched by this library so far  */
int mv_profile(mv_handle* h, int32_t enable);
int mv_profile_read(mv_handle* h, double* ms, double* flops, int32_t* count, int32_t n_tags);
                                             /* tags: 0 plain GEMM, 1 attn fwd, 2 attn bwd, 3 wgrad GEMM, 4 GELU-fwd GEMM, 5 GELU-bwd GEMM, */
                                             /* 6 residual-epilogue GEMM; arrays of n_tags (>= 3) entries                                 */

/* ---- data-parallel gradient exchange (one process per GPU) ---- */
int mv_comm_unique_id(uint8_t out[128]);
int mv_comm_init(mv_handle* h, const uint8_t id[128], int32_t rank, int32_t world);
int mv_comm_allreduce_f32(mv_handle* h, float* buf, int64_t count, void* stream);   /* scalars (label counts)       */
int mv_comm_sync(mv_handle* h, void* stream);

/* ---- per-op entry points (parity tests, integration of single kernels) ---- */
int mv_gemm(const mv_gemm_desc* d, int32_t precision, void* stream);
int mv_attn_mask_dump(const uint8_t* mode, const int32_t* t_len, int32_t B, int32_t A, int32_t L, uint8_t* out, void* stream);
int mv_mask_classify(const int64_t* mask, int32_t dims, int32_t B, int32_t A, int32_t L, uint8_t* mode, int32_t* t_len,
                     int32_t* mismatches, void* stream);
int mv_attention_fwd(int32_t B, int32_t L, int32_t heads, int32_t A, const uint8_t* mode, const int32_t* t_len,
                     const void* qkv, void* ctx, float* lse, float dropout_p, uint64_t seed, uint32_t site,
                     int32_t precision, void* stream);
int mv_attention_bwd(int32_t B, int32_t L, int32_t heads, int32_t A, const uint8_t* mode, const int32_t* t_len,
                     const void* qkv, const void* ctx, const float* lse, const void* dctx, void* dqkv, float* dq_acc,
                     float* delta, float dropout_p, uint64_t seed, uint32_t site, int32_t precision, void* stream);
int mv_layernorm_fwd(const void* x, void* y, const float* gamma, const float* beta, int32_t rows, int32_t H, float eps,
                     int32_t precision, void* stream);
int mv_layernorm_bwd(const void* dy, const void* x, const float* gamma, void* dx, float* dgamma, float* dbeta,
                     int32_t rows, int32_t H, float eps, int32_t precision, void* stream);
int mv_mlm_ce(const float* logits, int64_t ld, const int64_t* labels, int32_t n, int32_t V, void* dlogits, float gscale,
              float* loss_sum, int32_t* correct, float* row_lse, int32_t* row_argmax, int32_t precision, void* stream);
/* BatchNorm2d (+residual add) (+ReLU) of the frozen ResNet-50 trunk on channels-last [rows = B*H*W, C] activations:
 * replaces torch.nn.BatchNorm2d inside ImageEncoder_cnn (models/image.py:50-56; train-mode statistics per
 * models/train_origin.py:72).  workspace: mv_bn_workspace_floats(rows, C) floats. */
int64_t mv_bn_workspace_floats(int64_t rows, int32_t C);
int mv_bn_forward(const void* x, const void* resid, void* y, int64_t rows, int32_t C, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, float momentum, float eps, int32_t training, int32_t relu,
                  float* workspace, int64_t ws_floats, int32_t precision, void* stream);
/* uint8 image [B,3,H,W] -> (x/255 - mean)/std in the activation dtype, channels-last [B,H,W,3]: the device-side form of
 * get_transforms (data/helper.py:20-27), so a step ships uint8 pixels over PCIe.  mean3/std3 are HOST pointers. */
int mv_normalize_u8(const uint8_t* src, void* dst, int64_t B, int64_t hw, int32_t cpad, const float* mean3, const float* std3,
                    int32_t precision, void* stream);        /* dst is [B,H,W,cpad], channels >= 3 zero-filled        */
/* The same transform in 2x2 space-to-depth form, dst [B, H/2 + 3, W/2 + 3, 16] channels-last (zero border: 2 blocks in
 * front, 1 behind; channel = c*4 + dy*2 + dx as torch pixel_unshuffle, 12..15 zero): the ResNet stem's 7x7/2 convolution
 * (torchvision resnet50 child 0 inside ImageEncoder_cnn, models/image.py:50-56) is then a 4x4 stride-1 convolution over 16
 * channels with re-indexed weights — same arithmetic, a shape cuDNN runs on tensor cores.  H and W must be even. */
int mv_normalize_u8_s2d(const uint8_t* src, void* dst, int32_t B, int32_t H, int32_t W, const float* mean3, const float* std3,
                        int32_t precision, void* stream);
/* ResNet stem tail in one pass: BatchNorm + ReLU + MaxPool2d(3,2,1), channels-last [B,H,W,C] -> [B,H/2,W/2,C]
 * (torchvision resnet50 children 1-3 inside ImageEncoder_cnn, models/image.py:50-56). Workspace as mv_bn_forward. */
int mv_bn_relu_maxpool(const void* x, void* y, int32_t B, int32_t H, int32_t W, int32_t C, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, float momentum, float eps, int32_t training, float* workspace,
                       int64_t ws_floats, int32_t precision, void* stream);
/* ImageTextMatching.forward as a stand-alone op (models/cxrbert_origin.py:164-173; Downstream_task/Retrieval/retrieval.py:32):
 * logits[B,2] = pooled[B,H] . w[2,H]^T + b; with dlogits != NULL also the backward: d_pooled[B,H] (activation dtype),
 * dw[2,H] += , db[2] += (fp32). */
int mv_itm_head(const void* pooled, const float* w, const float* b, float* logits, int32_t B, int32_t H, const float* dlogits,
                void* d_pooled, float* dw, float* db, int32_t precision, void* stream);
/* Stem convolution of ImageEncoder_cnn (torchvision resnet50 conv1, 7x7 / stride 2 / pad 3, models/image.py:50-56) in its
 * space-to-depth form: x [B, Hs, Ws, 16] bf16 channels-last as written by mv_normalize_u8_s2d (2x2 pixel blocks, zero border),
 * w [O, 4, 4, 16] bf16 (the re-indexed 7x7 weights, one K-major row of 256 per output channel), y [B, Hs-3, Ws-3, O] bf16
 * channels-last.  Runs as a tcgen05 GEMM whose A rows are overlapping 64-element windows of x read in place through a 4-D
 * tensor map (no im2col copy).  (Ws - 3) % 128 == 0, O % 8 == 0. */
int mv_stem_conv_s2d(const void* x, const void* w, void* y, int32_t B, int32_t Hs, int32_t Ws, int32_t O, void* stream);
int mv_adamw(float* p, float* g, float* m, float* v, void* shadow_bf16, int64_t n, float lr, float beta1, float beta2,
             float eps, float weight_decay, int32_t step, float grad_scale, int32_t zero_grad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MEDVILL_SM100_H_ */
