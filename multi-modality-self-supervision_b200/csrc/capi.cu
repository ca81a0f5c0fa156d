// capi.cu — extern "C" surface of libmedvill_sm100.so (see include/medvill_sm100.h for the reference citations).
#include <string.h>

#include "engine.h"

namespace mv {
int nccl_unique_id(uint8_t out[128]);
int engine_comm_init(Engine* e, const uint8_t id[128], int rank, int world);
int engine_allreduce(Engine* e, float* buf, int64_t count, cudaStream_t s);
int engine_comm_sync(Engine* e, cudaStream_t s);
}  // namespace mv

using namespace mv;

struct mv_handle {
  Engine eng;
};

#define MV_CHECK_HANDLE(h) do { if (!(h)) { mv::set_error("null mv_handle"); return -1; } } while (0)
static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

const char* mv_last_error(void) { return mv::last_error(); }
int mv_abi_version(void) { return MV_ABI_VERSION; }

int mv_layout_query(const mv_config* cfg, mv_layout* out) {
  MV_REQUIRE(cfg && out, "mv_layout_query: null argument");
  return layout_compute(*cfg, out);
}

int mv_bucket_plan(const mv_config* cfg, int64_t* offsets, int64_t* counts, int32_t max_buckets, int32_t* n_buckets) {
  MV_REQUIRE(cfg && offsets && counts && n_buckets, "mv_bucket_plan: null argument");
  mv_layout lay;
  if (layout_compute(*cfg, &lay)) return -1;
  const std::vector<Bucket> b = bucket_plan(*cfg, lay);
  MV_REQUIRE(static_cast<int>(b.size()) <= max_buckets, "mv_bucket_plan: need room for %d buckets", (int)b.size());
  for (size_t i = 0; i < b.size(); ++i) { offsets[i] = b[i].offset; counts[i] = b[i].count; }
  *n_buckets = static_cast<int32_t>(b.size());
  return 0;
}

int mv_create(mv_handle** out, const mv_config* cfg) {
  MV_REQUIRE(out && cfg, "mv_create: null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    mv::set_error("no CUDA device visible: libmedvill_sm100 has no CPU fallback");
    return -4;
  }
  mv_handle* h = new mv_handle();
  const int rc = h->eng.init(*cfg);
  if (rc) { h->eng.destroy(); delete h; return rc; }
  *out = h;
  return 0;
}

int mv_destroy(mv_handle* h) {
  if (!h) return 0;
  h->eng.destroy();
  delete h;
  return 0;
}

int mv_bind_arenas(mv_handle* h, float* params, float* grads, float* adam_m, float* adam_v, void* shadow_bf16) {
  MV_CHECK_HANDLE(h);
  MV_REQUIRE(params && grads, "mv_bind_arenas: params and grads are required");
  MV_REQUIRE(h->eng.f32 || shadow_bf16, "mv_bind_arenas: bf16 precision needs the shadow arena");
  h->eng.params = params; h->eng.grads = grads; h->eng.adam_m = adam_m; h->eng.adam_v = adam_v;
  h->eng.shadow = static_cast<bf16*>(shadow_bf16);
  return 0;
}

int mv_refresh_shadow(mv_handle* h, void* stream) {
  MV_CHECK_HANDLE(h);
  if (!h->eng.shadow) return 0;
  return cast_f32_to_bf16(h->eng.params, h->eng.shadow, h->eng.lay.total, S(stream));
}

int mv_stats_reset(mv_handle* h, void* stream) {
  MV_CHECK_HANDLE(h);
  MV_CUDA_CHECK(cudaMemsetAsync(h->eng.stats, 0, sizeof(mv_step_stats), S(stream)));
  return 0;
}

int mv_forward(mv_handle* h, const mv_batch* b, void* stream) {
  MV_CHECK_HANDLE(h);
  MV_REQUIRE(b, "mv_forward: null batch");
  return h->eng.forward(*b, S(stream));
}

int mv_backward(mv_handle* h, const mv_batch* b, int32_t allreduce, void* stream) {
  MV_CHECK_HANDLE(h);
  MV_REQUIRE(b, "mv_backward: null batch");
  return h->eng.backward(*b, allreduce, S(stream));
}

int mv_backward_external(mv_handle* h, const mv_batch* b, const mv_external_grads* g, int32_t allreduce, void* stream) {
  MV_CHECK_HANDLE(h);
  MV_REQUIRE(b && g, "mv_backward_external: null argument");
  return h->eng.backward_external(*b, *g, allreduce, S(stream));
}

int mv_zero_grads(mv_handle* h, void* stream) {
  MV_CHECK_HANDLE(h);
  MV_REQUIRE(h->eng.grads, "arenas not bound");
  MV_CUDA_CHECK(cudaMemsetAsync(h->eng.grads, 0, h->eng.lay.total * sizeof(float), S(stream)));
  return 0;
}

int mv_adamw_step(mv_handle* h, float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                  float grad_scale, void* stream) {
  MV_CHECK_HANDLE(h);
  Engine& e = h->eng;
  MV_REQUIRE(e.params && e.grads && e.adam_m && e.adam_v, "mv_adamw_step: arenas (incl. Adam moments) not bound");
  AdamArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.step = step;
  a.grad_scale = grad_scale; a.zero_grad = 1;
  auto range = [&](int64_t off, int64_t n) {
    a.n = n; a.p = e.params + off; a.g = e.grads + off; a.m = e.adam_m + off; a.v = e.adam_v + off;
    a.shadow = e.shadow ? e.shadow + off : nullptr;
    return adamw_step(a, S(stream));
  };
  const int64_t head = e.lay.layer0;          // [0, layer0) = the embeddings bucket, reduced last (bucket_plan)
  if (e.comm_pending && e.mid_recorded && head % 4 == 0 && head > 0) {
    // the last bucket's all-reduce (word embeddings: ~100 MB) is still in flight when backward ends: update the layers
    // and heads behind the mid event while it finishes, then the embeddings
    MV_CUDA_CHECK(cudaStreamWaitEvent(S(stream), e.ev_mid, 0));
    if (range(head, e.lay.total - head)) return -2;
    if (engine_comm_sync(&e, S(stream))) return -2;
    return range(0, head);
  }
  if (engine_comm_sync(&e, S(stream))) return -2;
  return range(0, e.lay.total);
}

// BertAdam.step of the report-generation fine-tune (optimization.py:112-182): `lr` is the already scheduled rate
// (lr * warmup_linear(step / t_total, warmup), host side).  Waits for pending all-reduces; zeroes the updated gradients.
int mv_bert_adam_step(mv_handle* h, float lr, float beta1, float beta2, float eps, float weight_decay, float max_grad_norm,
                      void* stream) {
  MV_CHECK_HANDLE(h);
  if (engine_comm_sync(&h->eng, S(stream))) return -2;
  return h->eng.bert_adam(lr, beta1, beta2, eps, weight_decay, max_grad_norm, S(stream));
}

int mv_read_stats(mv_handle* h, mv_step_stats* host_out, void* stream) {
  MV_CHECK_HANDLE(h);
  MV_REQUIRE(host_out, "mv_read_stats: null output");
  MV_CUDA_CHECK(cudaMemcpyAsync(h->eng.stats_host, h->eng.stats, sizeof(mv_step_stats), cudaMemcpyDeviceToHost, S(stream)));
  MV_CUDA_CHECK(cudaStreamSynchronize(S(stream)));
  *host_out = *h->eng.stats_host;
  return 0;
}

// Enqueue the device->host copy of the step statistics into caller-owned PINNED memory; no synchronisation.  The caller
// orders its read with an event recorded on `stream` after this call (the next step's mv_stats_reset is stream-ordered
// behind the copy, so the values cannot be overwritten early).
int mv_read_stats_async(mv_handle* h, mv_step_stats* pinned_out, void* stream) {
  MV_CHECK_HANDLE(h);
  MV_REQUIRE(pinned_out, "mv_read_stats_async: null output");
  MV_CUDA_CHECK(cudaMemcpyAsync(pinned_out, h->eng.stats, sizeof(mv_step_stats), cudaMemcpyDeviceToHost, S(stream)));
  return 0;
}

int mv_itm_logits(mv_handle* h, float* host_out, int32_t B, void* stream) {
  MV_CHECK_HANDLE(h);
  MV_REQUIRE(host_out && B > 0 && B <= h->eng.cfg.max_batch, "mv_itm_logits: bad arguments");
  MV_CUDA_CHECK(cudaMemcpyAsync(host_out, h->eng.itm_logits, sizeof(float) * 2 * B, cudaMemcpyDeviceToHost, S(stream)));
  MV_CUDA_CHECK(cudaStreamSynchronize(S(stream)));
  return 0;
}

// Match probability softmax(itm_logits)[:, 1] of the last mv_forward, written to DEVICE memory (no synchronisation): the
// retrieval scorer appends one slice of the [images x reports] similarity matrix per forward.
int mv_itm_match_prob(mv_handle* h, float* device_out, int32_t B, void* stream) {
  MV_CHECK_HANDLE(h);
  MV_REQUIRE(device_out && B > 0 && B <= h->eng.cfg.max_batch, "mv_itm_match_prob: bad arguments");
  return itm_match_prob(h->eng.itm_logits, device_out, B, S(stream));
}

int mv_full_logits(mv_handle* h, const mv_batch* b, float* logits, int64_t ld, void* stream) {
  MV_CHECK_HANDLE(h);
  MV_REQUIRE(b && logits, "mv_full_logits: null argument");
  return h->eng.full_logits(*b, logits, ld, S(stream));
}

// Copy a named intermediate to `dst` (device or host pointer; cudaMemcpyDefault). Parity-test aid.
int mv_peek(mv_handle* h, const char* name, int32_t layer, void* dst, int64_t max_bytes, int64_t* bytes, void* stream) {
  MV_CHECK_HANDLE(h);
  MV_REQUIRE(name && dst && bytes, "mv_peek: null argument");
  Engine& e = h->eng;
  const size_t es = e.es, M = static_cast<size_t>(e.cfg.max_batch) * e.L, H = e.cfg.hidden, I = e.cfg.inter;
  const void* src = nullptr;
  size_t n = 0;
  const bool lay_ok = layer >= 0 && layer < e.cfg.layers;
  if (!strcmp(name, "x") && layer >= 0 && layer <= e.cfg.layers) { src = e.x[layer]; n = M * H * es; }
  else if (!strcmp(name, "emb_sum")) { src = e.emb_sum; n = M * H * es; }
  else if (!strcmp(name, "proj")) { src = e.proj; n = static_cast<size_t>(e.cfg.max_batch) * e.cfg.num_image_embeds * H * es; }
  else if (!strcmp(name, "qkv") && lay_ok) { src = e.lw[layer].qkv; n = M * 3 * H * es; }
  else if (!strcmp(name, "ctx") && lay_ok) { src = e.lw[layer].ctx; n = M * H * es; }
  else if (!strcmp(name, "x1") && lay_ok) { src = e.lw[layer].x1; n = M * H * es; }
  else if (!strcmp(name, "h1") && lay_ok) { src = e.lw[layer].h1; n = M * I * es; }
  else if (!strcmp(name, "lse") && lay_ok) { src = e.lw[layer].lse; n = static_cast<size_t>(e.cfg.max_batch) * e.nh * e.L * 4; }
  else if (!strcmp(name, "pooled")) { src = e.pooled; n = static_cast<size_t>(e.cfg.max_batch) * H * es; }
  else if (!strcmp(name, "itm_logits")) { src = e.itm_logits; n = static_cast<size_t>(e.cfg.max_batch) * 2 * 4; }
  else if (!strcmp(name, "logits")) { src = e.logits; n = static_cast<size_t>(e.mlm_cap) * e.Vpad * 4; }
  else if (!strcmp(name, "row_lse")) { src = e.row_lse; n = static_cast<size_t>(e.mlm_cap) * 4; }
  else if (!strcmp(name, "row_argmax")) { src = e.row_argmax; n = static_cast<size_t>(e.mlm_cap) * 4; }
  else if (!strcmp(name, "dx")) { src = e.dxa; n = M * H * es; }
  MV_REQUIRE(src != nullptr, "mv_peek: unknown buffer '%s' (layer %d)", name, layer);
  if (static_cast<int64_t>(n) > max_bytes) n = static_cast<size_t>(max_bytes);
  MV_CUDA_CHECK(cudaMemcpyAsync(dst, src, n, cudaMemcpyDefault, S(stream)));
  MV_CUDA_CHECK(cudaStreamSynchronize(S(stream)));
  *bytes = static_cast<int64_t>(n);
  return 0;
}

long mv_launch_count(void) { return mv::launch_count(); }

int mv_profile(mv_handle* h, int32_t enable) {
  MV_CHECK_HANDLE(h);
  h->eng.profiling = enable != 0;
  return 0;
}

// Sum the recorded event pairs per tag (engine.h: 0 / 3 / 4 / 5 / 6 GEMM classes, 1 attention fwd, 2 attention bwd); syncs,
// then clears the records.  The arrays hold n_tags entries (tags >= n_tags are folded into tag 0).
int mv_profile_read(mv_handle* h, double* ms, double* flops, int32_t* count, int32_t n_tags) {
  MV_CHECK_HANDLE(h);
  MV_REQUIRE(ms && flops && count && n_tags >= 3, "mv_profile_read: bad arguments");
  MV_CUDA_CHECK(cudaDeviceSynchronize());
  for (int i = 0; i < n_tags; ++i) { ms[i] = 0; flops[i] = 0; count[i] = 0; }
  for (auto& r : h->eng.prof) {
    float t = 0.f;
    const int tag = (r.tag >= 0 && r.tag < n_tags) ? r.tag : 0;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[tag] += t; flops[tag] += r.flops; count[tag] += 1; }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  h->eng.prof.clear();
  return 0;
}

int mv_comm_unique_id(uint8_t out[128]) { return nccl_unique_id(out); }
int mv_comm_init(mv_handle* h, const uint8_t id[128], int32_t rank, int32_t world) {
  MV_CHECK_HANDLE(h);
  MV_REQUIRE(world >= 1 && rank >= 0 && rank < world, "mv_comm_init: bad rank/world");
  return engine_comm_init(&h->eng, id, rank, world);
}
int mv_comm_allreduce_f32(mv_handle* h, float* buf, int64_t count, void* stream) {
  MV_CHECK_HANDLE(h);
  return engine_allreduce(&h->eng, buf, count, S(stream));
}
int mv_comm_sync(mv_handle* h, void* stream) {
  MV_CHECK_HANDLE(h);
  return engine_comm_sync(&h->eng, S(stream));
}

int mv_gemm(const mv_gemm_desc* g, int32_t precision, void* stream) {
  MV_REQUIRE(g, "mv_gemm: null descriptor");
  GemmDesc d;
  d.M = g->M; d.N = g->N; d.K = g->K;
  d.A = g->A; d.lda = g->lda; d.a_mn = g->a_mn;
  d.B = g->B; d.ldb = g->ldb; d.b_mn = g->b_mn;
  d.C = g->C; d.ldc = g->ldc; d.c_f32 = g->c_f32; d.accumulate = g->accumulate;
  d.C2 = g->C2; d.ldc2 = g->ldc2;
  d.epi = g->epi; d.bias = g->bias; d.resid = g->resid; d.ldr = g->ldr; d.aux = g->aux; d.ldaux = g->ldaux; d.resid_f32 = g->resid_f32;
  d.drop_on = g->dropout_p > 0.f; d.drop_site = g->dropout_site; d.drop = make_dropout(g->dropout_p, g->dropout_seed);
  return precision == MV_PREC_FP32 ? gemm_f32_simt(d, S(stream)) : gemm_bf16_tc05(d, S(stream));
}

int mv_attn_mask_dump(const uint8_t* mode, const int32_t* t_len, int32_t B, int32_t A, int32_t L, uint8_t* out, void* stream) {
  MV_REQUIRE(mode && t_len && out, "mv_attn_mask_dump: null argument");
  return mask_dump(mode, t_len, B, A, L, out, S(stream));
}

int mv_mask_classify(const int64_t* mask, int32_t dims, int32_t B, int32_t A, int32_t L, uint8_t* mode, int32_t* t_len,
                     int32_t* mismatches, void* stream) {
  MV_REQUIRE(mask && mode && t_len && mismatches, "mv_mask_classify: null argument");
  return mask_classify(reinterpret_cast<const int64_t*>(mask), dims, B, A, L, mode, t_len, mismatches, S(stream));
}

static AttnArgs make_attn(int32_t B, int32_t L, int32_t heads, int32_t A, const uint8_t* mode, const int32_t* t_len,
                          const void* qkv, void* ctx, float* lse, float p, uint64_t seed, uint32_t site) {
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.L = L; a.nh = heads; a.A = A; a.mode = mode; a.t_len = t_len; a.qkv = qkv; a.ctx = ctx; a.lse = lse;
  a.drop_on = p > 0.f; a.drop_site = site; a.drop = make_dropout(p, seed);
  return a;
}

int mv_attention_fwd(int32_t B, int32_t L, int32_t heads, int32_t A, const uint8_t* mode, const int32_t* t_len,
                     const void* qkv, void* ctx, float* lse, float dropout_p, uint64_t seed, uint32_t site,
                     int32_t precision, void* stream) {
  AttnArgs a = make_attn(B, L, heads, A, mode, t_len, qkv, ctx, lse, dropout_p, seed, site);
  return precision == MV_PREC_FP32 ? attention_fwd_simt(a, S(stream)) : attention_fwd_tc05(a, S(stream));
}

int mv_attention_bwd(int32_t B, int32_t L, int32_t heads, int32_t A, const uint8_t* mode, const int32_t* t_len,
                     const void* qkv, const void* ctx, const float* lse, const void* dctx, void* dqkv, float* dq_acc,
                     float* delta, float dropout_p, uint64_t seed, uint32_t site, int32_t precision, void* stream) {
  AttnArgs a = make_attn(B, L, heads, A, mode, t_len, qkv, const_cast<void*>(ctx), const_cast<float*>(lse), dropout_p, seed, site);
  a.dctx = dctx; a.dqkv = dqkv; a.dq_acc = dq_acc; a.delta = delta;
  return precision == MV_PREC_FP32 ? attention_bwd_simt(a, S(stream)) : attention_bwd_tc05(a, S(stream));
}

int mv_layernorm_fwd(const void* x, void* y, const float* gamma, const float* beta, int32_t rows, int32_t H, float eps,
                     int32_t precision, void* stream) {
  return ln_fwd(x, y, gamma, beta, rows, H, eps, 0, 0, make_dropout(0.f, 0), precision == MV_PREC_FP32, S(stream));
}

int mv_layernorm_bwd(const void* dy, const void* x, const float* gamma, void* dx, float* dgamma, float* dbeta,
                     int32_t rows, int32_t H, float eps, int32_t precision, void* stream) {
  return ln_bwd(dy, x, gamma, dx, nullptr, dgamma, dbeta, nullptr, rows, H, eps, 0, 0, 0, make_dropout(0.f, 0),
                precision == MV_PREC_FP32, S(stream));
}

int mv_mlm_ce(const float* logits, int64_t ld, const int64_t* labels, int32_t n, int32_t V, void* dlogits, float gscale,
              float* loss_sum, int32_t* correct, float* row_lse, int32_t* row_argmax, int32_t precision, void* stream) {
  CeArgs a;
  a.n = n; a.V = V; a.ldv = ld; a.logits = logits; a.labels = reinterpret_cast<const int64_t*>(labels); a.dlogits = dlogits;
  a.gscale = gscale; a.loss_sum = loss_sum; a.correct = correct; a.row_lse = row_lse; a.row_argmax = row_argmax;
  return mlm_ce_fwd_bwd(a, precision == MV_PREC_FP32, S(stream));
}

int64_t mv_bn_workspace_floats(int64_t rows, int32_t C) { return bn_workspace_floats(rows, C); }

int mv_bn_forward(const void* x, const void* resid, void* y, int64_t rows, int32_t C, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, float momentum, float eps, int32_t training, int32_t relu,
                  float* workspace, int64_t ws_floats, int32_t precision, void* stream) {
  return bn_forward(x, resid, y, rows, C, gamma, beta, running_mean, running_var, momentum, eps, training, relu, workspace,
                    ws_floats, precision == MV_PREC_FP32, S(stream));
}

int mv_normalize_u8(const uint8_t* src, void* dst, int64_t B, int64_t hw, int32_t cpad, const float* mean3, const float* std3,
                    int32_t precision, void* stream) {
  MV_REQUIRE(mean3 && std3, "mv_normalize_u8: null mean/std");
  return normalize_u8(src, dst, B, hw, cpad, mean3, std3, precision == MV_PREC_FP32, S(stream));
}

int mv_normalize_u8_s2d(const uint8_t* src, void* dst, int32_t B, int32_t H, int32_t W, const float* mean3, const float* std3,
                        int32_t precision, void* stream) {
  MV_REQUIRE(mean3 && std3, "mv_normalize_u8_s2d: null mean/std");
  return normalize_u8_s2d(src, dst, B, H, W, mean3, std3, precision == MV_PREC_FP32, S(stream));
}

int mv_bn_relu_maxpool(const void* x, void* y, int32_t B, int32_t H, int32_t W, int32_t C, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, float momentum, float eps, int32_t training, float* workspace,
                       int64_t ws_floats, int32_t precision, void* stream) {
  return bn_relu_maxpool(x, y, B, H, W, C, gamma, beta, running_mean, running_var, momentum, eps, training, workspace, ws_floats,
                         precision == MV_PREC_FP32, S(stream));
}

int mv_itm_head(const void* pooled, const float* w, const float* b, float* logits, int32_t B, int32_t H, const float* dlogits,
                void* d_pooled, float* dw, float* db, int32_t precision, void* stream) {
  MV_REQUIRE(pooled && w && b && logits && B > 0 && H > 0, "mv_itm_head: bad arguments");
  MV_REQUIRE(!dlogits || (d_pooled && dw && db), "mv_itm_head: backward needs d_pooled, dw, db");
  static float* scratch = nullptr;            // loss / accuracy sinks the fused kernel always writes
  if (!scratch) MV_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&scratch), 16));
  ItmArgs a;
  a.B = B; a.H = H; a.pooled = pooled; a.w = w; a.b = b; a.labels = nullptr; a.gscale = 1.f; a.logits = logits;
  a.loss_sum = scratch; a.correct = reinterpret_cast<int*>(scratch + 1);
  a.d_pre = dlogits ? d_pooled : nullptr; a.dw = dw; a.db = db; a.ext_dlogits = dlogits; a.no_tanh = 1;
  return itm_head_fwd_bwd(a, precision == MV_PREC_FP32, S(stream));
}

int mv_stem_conv_s2d(const void* x, const void* w, void* y, int32_t B, int32_t Hs, int32_t Ws, int32_t O, void* stream) {
  MV_REQUIRE(x && w && y && B > 0 && Hs > 3 && Ws > 3 && O > 0 && O % 8 == 0, "mv_stem_conv_s2d: bad arguments");
  MV_REQUIRE((Ws - 3) % 128 == 0, "mv_stem_conv_s2d: output width %d must be a multiple of 128", Ws - 3);
  GemmDesc d;
  d.M = B * (Hs - 3) * (Ws - 3); d.N = O; d.K = 256;
  d.A = x; d.lda = 64; d.a_mn = 0;
  d.B = w; d.ldb = 256; d.b_mn = 0;
  d.C = y; d.ldc = O; d.c_f32 = 0;
  d.epi = EPI_NONE;
  d.a_win = 1; d.win_b = B; d.win_ho = Hs - 3; d.win_wo = Ws - 3; d.win_hs = Hs; d.win_ws = Ws; d.win_c = 16;
  return gemm_bf16_tc05(d, S(stream));
}

int mv_adamw(float* p, float* g, float* m, float* v, void* shadow_bf16, int64_t n, float lr, float beta1, float beta2,
             float eps, float weight_decay, int32_t step, float grad_scale, int32_t zero_grad, void* stream) {
  AdamArgs a;
  a.n = n; a.p = p; a.g = g; a.m = m; a.v = v; a.shadow = static_cast<bf16*>(shadow_bf16);
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.step = step;
  a.grad_scale = grad_scale; a.zero_grad = zero_grad;
  return adamw_step(a, S(stream));
}

}  // extern "C"
