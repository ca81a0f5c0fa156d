// engine.h — the per-GPU engine behind mv_handle: flat parameter arena layout, activation workspaces, the
// forward / backward / optimizer drivers of the MedViLL pre-training step, and the NCCL gradient exchange.
#pragma once
#include <vector>

#include "../../include/medvill_sm100.h"
#include "gemm.h"
#include "kernels.h"

namespace mv {

int layout_compute(const mv_config& c, mv_layout* out);

struct Bucket { int64_t offset, count; };
// gradient buckets in the order backward finishes them: heads tail block, layers L-1..0, embeddings head block
std::vector<Bucket> bucket_plan(const mv_config& c, const mv_layout& lay);

struct LayerWs {
  void *qkv, *ctx, *y1, *x1, *h1, *g1, *y2;
  float* lse;
};

struct NcclApi;

struct Engine {
  mv_config cfg;
  mv_layout lay;
  int f32 = 0;          // activation dtype is fp32 (check mode)
  size_t es = 2;        // bytes per activation element
  int L = 0, A = 0, T = 0, nh = 0, Vpad = 0;
  int rf32 = 0;         // bf16 mode with the residual stream (pre-LN sums y1 / y2 and the LN outputs that feed the next residual add) in fp32

  // borrowed arenas
  float* params = nullptr; float* grads = nullptr; float* adam_m = nullptr; float* adam_v = nullptr; bf16* shadow = nullptr;

  // owned workspaces
  std::vector<void*> allocs;
  std::vector<void*> x;          // [layers + 1] : x[0] = embedding output, x[l+1] = output of layer l
  std::vector<LayerWs> lw;
  void *emb_sum = nullptr, *proj = nullptr, *feats_g = nullptr;
  float *xres = nullptr, *x1res = nullptr;   // rf32: fp32 copies of x[l] / x1 (transient: consumed by the next residual epilogue)
  void *cls_rows = nullptr, *pooled = nullptr, *d_pre = nullptr, *d_cls = nullptr;
  float* itm_logits = nullptr;
  void *dxa = nullptr, *dxb = nullptr, *dxc = nullptr, *dh1 = nullptr, *dqkv = nullptr, *dctx = nullptr, *dproj = nullptr;
  float *dq_acc = nullptr, *delta = nullptr;
  float* dq_part = nullptr;        // MV_FLAG_DETERMINISTIC: per-key-tile dQ partials [ceil(L/128)][B][L][H]
  int64_t* zero_idx = nullptr;     // [1] = {0}: gather/scatter of the [CLS] rows
  mv_step_stats* stats = nullptr;    // device
  mv_step_stats* stats_host = nullptr;  // pinned
  // MLM head workspace (grown on demand)
  int mlm_cap = 0;
  void *rows_h = nullptr, *t_pre = nullptr, *t_act = nullptr, *t_ln = nullptr, *dlogits = nullptr, *d_tln = nullptr,
       *d_tact = nullptr, *d_tpre = nullptr, *d_rows = nullptr;
  float* logits = nullptr; float* row_lse = nullptr; int* row_argmax = nullptr;
  float* row_aux = nullptr;          // drop-worst: row_loss[cap] | row_scale[cap] | scratch {float loss, int correct}

  // data-parallel state
  NcclApi* nccl = nullptr; void* comm = nullptr; int rank = 0, world = 1;
  cudaStream_t comm_stream = nullptr; cudaEvent_t ev_ready = nullptr, ev_done = nullptr; bool comm_pending = false;
  cudaEvent_t ev_mid = nullptr; bool mid_recorded = false;   // all buckets but the last (embeddings) have been reduced
  std::vector<Bucket> buckets;
  bf16* comm_stage = nullptr;      // bf16 gradient exchange (MEDVILL_GRAD_COMM=bf16): staging for the largest bucket
  int comm_bf16 = 0;

  // optional per-kernel-family timing (CUDA events on the launch stream): tag 0 = tcgen05/SIMT GEMM, 1 = attention fwd,
  // 2 = attention bwd.  Used by bench.py for the live roofline numbers; off by default.
  struct ProfRec { cudaEvent_t a, b; int tag; double flops; };
  bool profiling = false;
  std::vector<ProfRec> prof;
  int prof_begin(int tag, double flops, cudaStream_t s);
  int prof_end(cudaStream_t s);

  // BertAdam state (built on first use): chunk table + per-tensor norm scratch, device
  AdamChunk* adam_chunks = nullptr; float* adam_sumsq = nullptr; int n_adam_chunks = 0, n_adam_tensors = 0;
  int bert_adam(float lr, float beta1, float beta2, float eps, float weight_decay, float max_grad_norm, cudaStream_t s);

  int init(const mv_config& c);
  void destroy();
  int ensure_mlm(int n);
  int alloc(void** p, size_t bytes);

  const void* W(int64_t off) const { return f32 ? static_cast<const void*>(params + off) : static_cast<const void*>(shadow + off); }
  int gemm(const GemmDesc& d, cudaStream_t s) {
    // tags: 0 plain forward / dgrad GEMM, 3 weight-gradient GEMM (fp32 reduce-add), 4 GELU-forward epilogue,
    //       5 GELU-backward epilogue, 6 residual epilogue (fp32 stream); 1 / 2 are the attention kernels
    const int tag = d.accumulate ? 3 : ((d.epi == EPI_BIAS_GELU || d.epi == EPI_BIAS_GELU_GRAD) ? 4 : ((d.epi == EPI_DGELU || d.epi == EPI_MUL) ? 5 :
                    ((d.epi == EPI_BIAS_RESID || d.epi == EPI_RESID) ? 6 : 0)));
    if (prof_begin(tag, 2.0 * d.M * d.N * d.K, s)) return -2;
    const int rc = f32 ? gemm_f32_simt(d, s) : gemm_bf16_tc05(d, s);
    if (rc) return rc;
    return prof_end(s);
  }
  int linear_fwd(const void* X, int M, int K, int64_t w_off, int N, int64_t b_off, void* Y, int epi, void* pre,
                 const void* resid, int drop_on, uint32_t site, const DropoutCfg& dc, cudaStream_t s, int y_f32 = 0, long ldy = 0,
                 int resid_f32 = 0);
  int linear_dgrad(const void* dY, long lddy, int M, int N, int64_t w_off, int K, void* dX, int epi, const void* extra, cudaStream_t s);
  int linear_wgrad(const void* dY, long lddy, const void* X, int M, int N, int K, int64_t w_off, cudaStream_t s);

  int forward(const mv_batch& b, cudaStream_t s);
  int backward(const mv_batch& b, int allreduce, cudaStream_t s);
  int full_logits(const mv_batch& b, float* out, int64_t ld, cudaStream_t s);
  int mlm_head_rows(const int64_t* rows, int n, const DropoutCfg& dc, cudaStream_t s);   // rows of x[layers] -> rows_h, t_pre, t_act, t_ln
  int backward_external(const mv_batch& b, const mv_external_grads& g, int allreduce, cudaStream_t s);
  const void* ext_dlogits = nullptr;     // backward(): gradient of the labelled-row logits supplied by the caller (autograd path)
  const void* ext_dseq = nullptr;        // backward(): gradient w.r.t. the final hidden states [B*L, H] supplied by the caller
  int bucket_done(size_t idx, int allreduce, cudaStream_t s);
};

}  // namespace mv
