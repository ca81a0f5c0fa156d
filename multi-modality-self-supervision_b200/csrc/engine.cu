// engine.cu — drivers of the MedViLL pre-training step on one B200.
//
// Data layout in HBM (B = micro-batch, L = N + S + 3 joint length, M = B*L rows, H hidden, I intermediate):
//   * one flat fp32 parameter arena (+ identical-layout grad / Adam m / Adam v arenas and a bf16 shadow used as GEMM
//     operands); nn.Linear weights stay [out, in] so forward is a K-major x K-major ("TN") GEMM, dgrad reads the same
//     weight MN-major and wgrad reads both activations MN-major — nothing is ever transposed in memory;
//   * activations are row-major [M, features] in the activation dtype; Q|K|V share one [M, 3H] buffer that the
//     attention kernels read with strided TMA boxes; everything backward needs is kept per layer
//     (qkv, ctx, pre-LN sums y1/y2, x1, pre-GELU h1, GELU g1, row LSE) — 686 MB per layer at B=64 in bf16;
//   * dropout masks are never stored: regenerated from (step seed, site id, element index).
//
// Reference call stack replaced: CXRBERT.forward (models/cxrbert_origin.py:144-149) -> CXRBertEncoder.forward (:87-130)
// -> upstream BertEncoder x12 -> heads (:205-248, :164-173) -> CE losses (models/train_origin.py:118-126) -> backward
// -> AdamW (:129-131), with nn.DataParallel's gradient reduction (:53-55) replaced by bucketed NCCL all-reduce.
#include "engine.h"

#include <dlfcn.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace mv {

namespace {
inline int64_t up64(int64_t x) { return (x + 63) / 64 * 64; }

enum : uint32_t { SITE_EMB = 1, SITE_LAYER0 = 16 };
inline uint32_t site_att(int l) { return SITE_LAYER0 + 4 * l; }
inline uint32_t site_h1(int l) { return SITE_LAYER0 + 4 * l + 1; }
inline uint32_t site_h2(int l) { return SITE_LAYER0 + 4 * l + 2; }
}  // namespace

int layout_compute(const mv_config& c, mv_layout* o) {
  MV_REQUIRE(c.hidden > 0 && c.hidden % 64 == 0 && c.hidden <= 1024, "hidden must be a multiple of 64, <= 1024 (got %d)", c.hidden);
  MV_REQUIRE(c.heads * 64 == c.hidden, "head dim must be 64: hidden=%d heads=%d", c.hidden, c.heads);
  MV_REQUIRE(c.inter > 0 && c.inter % 64 == 0, "intermediate size must be a multiple of 64 (got %d)", c.inter);
  MV_REQUIRE(c.img_hidden > 0 && c.img_hidden % 64 == 0, "img_hidden must be a multiple of 64 (got %d)", c.img_hidden);
  MV_REQUIRE(c.layers > 0 && c.vocab > 0 && c.max_pos > 0 && c.type_vocab > 0, "bad BERT dims");
  MV_REQUIRE(c.num_image_embeds > 0 && c.seq_len > 0 && c.grid >= c.num_image_embeds, "bad sequence dims (N=%d S=%d grid=%d)",
             c.num_image_embeds, c.seq_len, c.grid);
  MV_REQUIRE(c.seq_len + 1 <= c.max_pos && c.grid <= c.max_pos, "position table too small");
  const int64_t H = c.hidden, I = c.inter, V = c.vocab;
  int64_t off = 0;
  auto take = [&](int64_t n) { const int64_t at = off; off += up64(n); return at; };
  o->word = take(V * H); o->pos = take(static_cast<int64_t>(c.max_pos) * H); o->type = take(static_cast<int64_t>(c.type_vocab) * H);
  o->emb_ln_g = take(H); o->emb_ln_b = take(H);
  o->img_w = take(H * c.img_hidden); o->img_b = take(H);
  o->layer0 = off;
  int64_t r = 0;
  auto rel = [&](int64_t n) { const int64_t at = r; r += n; return at; };
  o->l_wqkv = rel(3 * H * H); o->l_bqkv = rel(3 * H); o->l_wo = rel(H * H); o->l_bo = rel(H);
  o->l_ln1_g = rel(H); o->l_ln1_b = rel(H); o->l_w1 = rel(I * H); o->l_b1 = rel(I); o->l_w2 = rel(H * I); o->l_b2 = rel(H);
  o->l_ln2_g = rel(H); o->l_ln2_b = rel(H);
  o->layer_stride = up64(r);
  off += o->layer_stride * c.layers;
  o->pool_w = take(H * H); o->pool_b = take(H);
  o->mlm_bias = take(V); o->mlm_tw = take(H * H); o->mlm_tb = take(H); o->mlm_ln_g = take(H); o->mlm_ln_b = take(H);
  o->itm_w = take(2 * H); o->itm_b = take(2);
  o->total = off;
  o->vocab_padded = up64(V);
  return 0;
}

std::vector<Bucket> bucket_plan(const mv_config& c, const mv_layout& lay) {
  std::vector<Bucket> b;
  b.push_back({lay.pool_w, lay.total - lay.pool_w});                       // heads (ready first)
  for (int l = c.layers - 1; l >= 0; --l) b.push_back({lay.layer0 + l * lay.layer_stride, lay.layer_stride});
  b.push_back({0, lay.layer0});                                            // embeddings + image projection (last)
  return b;
}

// ---- NCCL through dlopen: whichever libnccl.so.2 the process already mapped (torch's) is reused ----
struct NcclUid { char b[128]; };  // ncclUniqueId is passed by value
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
  int (*CommInitRankConfig)(void**, int, NcclUid, int, void*) = nullptr;   // optional (NCCL >= 2.14)
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

static NcclApi* load_nccl() {
  static NcclApi api;
  if (api.lib) return &api;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { set_error("cannot dlopen libnccl.so.2: %s", dlerror()); return nullptr; }
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
  api.CommInitRankConfig = reinterpret_cast<decltype(api.CommInitRankConfig)>(dlsym(lib, "ncclCommInitRankConfig"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(lib, "ncclAllReduce"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) {
    set_error("libnccl.so.2 lacks required symbols");
    return nullptr;
  }
  api.lib = lib;
  return &api;
}

int nccl_unique_id(uint8_t out[128]) {
  NcclApi* n = load_nccl();
  if (!n) return -3;
  const int rc = n->GetUniqueId(out);
  MV_REQUIRE(rc == 0, "ncclGetUniqueId failed: %s", n->GetErrorString ? n->GetErrorString(rc) : "?");
  return 0;
}

// SMs the persistent GEMMs leave free for NCCL's CTAs while gradient buckets are in flight.  Without this the all-reduce
// kernel's CTAs queue behind a 148-CTA persistent GEMM, take over some SMs when it exits, and the NEXT GEMM's statically
// assigned tiles on those SMs wait for the whole collective (GEMM time per step 13.6 -> 15.6 ms at 8 GPUs; 14.8 ms with
// the reservation, profiles/r01_bench_n8.json).  447 MB of fp32 gradients per step need ~2 ms of NVLink time spread over
// a 25 ms backward, so a few CTAs are plenty.  MEDVILL_COMM_CTAS=n sets the reservation AND caps this communicator at n
// CTAs through ncclConfig_t.maxCTAs.  Unset: up to 4 ranks the reservation is 8 and NCCL keeps its own channel count (the
// configuration measured at 2 / 4 GPUs); from 8 ranks on both are 16 — the 8-GPU sweep profiles/r02_ctas_sweep_n8.jsonl:
// cap = reservation 8 / 12 / 16 -> 0.934 / 0.947 / 0.954 of 8 x 1 GPU, uncapped with 8 reserved 0.936 (uncapped, NCCL's
// CTAs spill onto the GEMMs' SMs: GEMM time 14.2 ms per step instead of 13.5; capped lower, the last bucket's tail grows).
static int comm_ctas(int world, bool* explicit_cap = nullptr) {
  const char* e = getenv("MEDVILL_COMM_CTAS");
  const bool from_env = e != nullptr && *e != 0;
  int n = from_env ? atoi(e) : (world >= 8 ? 16 : 8);
  if (n < 1) n = 1;
  if (n > 32) n = 32;
  if (explicit_cap) *explicit_cap = from_env || world >= 8;
  return n;
}

// The prefix of ncclConfig_t that has been stable since NCCL 2.17 (nccl.h: ncclConfig_v21700).  NCCL copies `size` bytes,
// checks the magic, and fills every field newer than `version` with its default, so this prefix is accepted by 2.17+.
struct NcclConfigV21700 {
  size_t size;
  unsigned int magic;
  unsigned int version;
  int blocking;
  int cgaClusterSize;
  int minCTAs;
  int maxCTAs;
  const char* netName;
};

int engine_comm_init(Engine* e, const uint8_t id[128], int rank, int world) {
  NcclApi* n = load_nccl();
  if (!n) return -3;
  NcclUid uid;
  memcpy(uid.b, id, 128);
  void* comm = nullptr;
  bool cap = false;
  const int ctas = comm_ctas(world, &cap);
  int rc;
  if (cap && n->CommInitRankConfig) {
    constexpr int kUndef = INT_MIN;                                  // NCCL_CONFIG_UNDEF_INT
    NcclConfigV21700 cfg{sizeof(NcclConfigV21700), 0xcafebeefu, 21700u, kUndef, kUndef, kUndef, ctas, nullptr};
    rc = n->CommInitRankConfig(&comm, world, uid, rank, &cfg);
    MV_REQUIRE(rc == 0, "ncclCommInitRankConfig(maxCTAs=%d) failed: %s", ctas, n->GetErrorString ? n->GetErrorString(rc) : "?");
  } else {
    rc = n->CommInitRank(&comm, world, uid, rank);
    MV_REQUIRE(rc == 0, "ncclCommInitRank failed: %s", n->GetErrorString ? n->GetErrorString(rc) : "?");
  }
  e->nccl = n; e->comm = comm; e->rank = rank; e->world = world;
  const char* gc = getenv("MEDVILL_GRAD_COMM");
  e->comm_bf16 = (!e->f32 && gc && !strcmp(gc, "bf16")) ? 1 : 0;
  if (e->comm_bf16 && !e->comm_stage) {
    int64_t mx = 0;
    for (const Bucket& b : e->buckets) mx = b.count > mx ? b.count : mx;
    void* p = nullptr;
    if (e->alloc(&p, static_cast<size_t>(mx) * sizeof(bf16))) return -2;
    e->comm_stage = static_cast<bf16*>(p);
  }
  MV_CUDA_CHECK(cudaStreamCreateWithFlags(&e->comm_stream, cudaStreamNonBlocking));
  MV_CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_ready, cudaEventDisableTiming));
  MV_CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_done, cudaEventDisableTiming));
  MV_CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_mid, cudaEventDisableTiming));
  return 0;
}

int engine_allreduce(Engine* e, float* buf, int64_t count, cudaStream_t s) {
  MV_REQUIRE(e->comm != nullptr, "communicator not initialised (mv_comm_init)");
  const int rc = e->nccl->AllReduce(buf, buf, static_cast<size_t>(count), /*ncclFloat32*/ 7, /*ncclSum*/ 0, e->comm, s);
  MV_REQUIRE(rc == 0, "ncclAllReduce failed: %s", e->nccl->GetErrorString ? e->nccl->GetErrorString(rc) : "?");
  return 0;
}

// Gradient bucket exchange.  MEDVILL_GRAD_COMM=bf16 (bf16 precision mode only): the bucket travels as bf16 — half the NVLink
// bytes and half the time the collective's CTAs share the SMs with the persistent GEMMs — and comes back into the fp32
// arena; the arena itself (accumulation across micro-batches, Adam) stays fp32.  Default: fp32 on the wire.
static int engine_allreduce_bucket(Engine* e, float* buf, int64_t count, cudaStream_t s) {
  if (!e->comm_bf16) return engine_allreduce(e, buf, count, s);
  MV_REQUIRE(e->comm != nullptr && e->comm_stage != nullptr, "communicator not initialised (mv_comm_init)");
  if (cast_f32_to_bf16(buf, e->comm_stage, count, s)) return -2;
  const int rc = e->nccl->AllReduce(e->comm_stage, e->comm_stage, static_cast<size_t>(count), /*ncclBfloat16*/ 9, /*ncclSum*/ 0, e->comm, s);
  MV_REQUIRE(rc == 0, "ncclAllReduce(bf16) failed: %s", e->nccl->GetErrorString ? e->nccl->GetErrorString(rc) : "?");
  return cast_bf16_to_f32(e->comm_stage, buf, count, s);
}

int engine_comm_sync(Engine* e, cudaStream_t s) {
  set_reserved_sms(0);
  if (e->comm_pending) {
    MV_CUDA_CHECK(cudaEventRecord(e->ev_done, e->comm_stream));
    MV_CUDA_CHECK(cudaStreamWaitEvent(s, e->ev_done, 0));
    e->comm_pending = false;
    e->mid_recorded = false;
  }
  return 0;
}

// --------------------------------------------------------------------------------------------------------------------
int Engine::prof_begin(int tag, double flops, cudaStream_t s) {
  if (!profiling) return 0;
  ProfRec r;
  r.tag = tag; r.flops = flops;
  MV_CUDA_CHECK(cudaEventCreate(&r.a));
  MV_CUDA_CHECK(cudaEventCreate(&r.b));
  MV_CUDA_CHECK(cudaEventRecord(r.a, s));
  prof.push_back(r);
  return 0;
}
int Engine::prof_end(cudaStream_t s) {
  if (!profiling) return 0;
  MV_CUDA_CHECK(cudaEventRecord(prof.back().b, s));
  return 0;
}

int Engine::alloc(void** p, size_t bytes) {
  *p = nullptr;
  if (bytes == 0) bytes = 16;
  MV_CUDA_CHECK(cudaMalloc(p, (bytes + 255) / 256 * 256));
  allocs.push_back(*p);
  return 0;
}

int Engine::init(const mv_config& c) {
  cfg = c;
  if (layout_compute(c, &lay)) return -1;
  MV_REQUIRE(c.max_batch > 0, "max_batch must be positive");
  MV_REQUIRE(c.precision == MV_PREC_BF16 || c.precision == MV_PREC_FP32, "unknown precision %d", c.precision);
  MV_REQUIRE(c.dropout_p >= 0.f && c.dropout_p < 1.f && c.attn_dropout_p >= 0.f && c.attn_dropout_p < 1.f && c.img_dropout_p >= 0.f &&
             c.img_dropout_p < 1.f, "dropout probabilities out of range");
  int dev = 0, major = 0;
  MV_CUDA_CHECK(cudaGetDevice(&dev));
  MV_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  MV_REQUIRE(major == 10, "libmedvill_sm100 needs an sm_100 (Blackwell B200) GPU; device %d is sm_%d0 — no fallback path", dev, major);
  f32 = c.precision == MV_PREC_FP32;
  es = f32 ? 4 : 2;
  rf32 = (!f32 && !(c.flags & MV_FLAG_BF16_RESIDUAL)) ? 1 : 0;
  nh = c.heads;
  A = c.num_image_embeds + 2; T = c.seq_len + 1; L = A + T;
  Vpad = static_cast<int>(lay.vocab_padded);
  const size_t M = static_cast<size_t>(c.max_batch) * L, H = c.hidden, I = c.inter, B = c.max_batch, N = c.num_image_embeds;
  const size_t ys = rf32 ? 4 : es;          // element size of the pre-LN sums
  x.resize(c.layers + 1);
  for (auto& p : x) if (alloc(&p, M * H * es)) return -2;
  lw.resize(c.layers);
  for (auto& w : lw) {
    if (alloc(&w.qkv, M * 3 * H * es) || alloc(&w.ctx, M * H * es) || alloc(&w.y1, M * H * ys) || alloc(&w.x1, M * H * es) ||
        alloc(&w.h1, M * I * es) || alloc(&w.g1, M * I * es) || alloc(&w.y2, M * H * ys) ||
        alloc(reinterpret_cast<void**>(&w.lse), B * nh * L * sizeof(float)))
      return -2;
  }
  if (alloc(&emb_sum, M * H * es) || alloc(&proj, B * N * H * es) || alloc(&feats_g, B * N * c.img_hidden * es) ||
      alloc(&cls_rows, B * H * es) || alloc(&pooled, B * H * es) || alloc(&d_pre, B * H * es) || alloc(&d_cls, B * H * es) ||
      alloc(reinterpret_cast<void**>(&itm_logits), B * 2 * sizeof(float)) || alloc(&dxa, M * H * es) || alloc(&dxb, M * H * es) ||
      alloc(&dxc, M * H * es) || alloc(&dh1, M * I * es) || alloc(&dqkv, M * 3 * H * es) || alloc(&dctx, M * H * es) ||
      alloc(&dproj, B * N * H * es) || alloc(reinterpret_cast<void**>(&dq_acc), M * H * sizeof(float)) ||
      alloc(reinterpret_cast<void**>(&delta), B * nh * L * sizeof(float)) || alloc(reinterpret_cast<void**>(&zero_idx), 16) ||
      alloc(reinterpret_cast<void**>(&stats), sizeof(mv_step_stats)))
    return -2;
  if (rf32 && (alloc(reinterpret_cast<void**>(&xres), M * H * 4) || alloc(reinterpret_cast<void**>(&x1res), M * H * 4))) return -2;
  if (!f32 && (c.flags & MV_FLAG_DETERMINISTIC) && alloc(reinterpret_cast<void**>(&dq_part), static_cast<size_t>((L + 127) / 128) * M * H * 4)) return -2;
  MV_CUDA_CHECK(cudaMemset(zero_idx, 0, 16));
  MV_CUDA_CHECK(cudaMemset(stats, 0, sizeof(mv_step_stats)));
  MV_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&stats_host), sizeof(mv_step_stats)));
  buckets = bucket_plan(cfg, lay);
  return ensure_mlm(static_cast<int>(B * T / 4 + 64));
}

void Engine::destroy() {
  cudaDeviceSynchronize();
  for (void* p : allocs) cudaFree(p);
  allocs.clear();
  void* mlm[] = {rows_h, t_pre, t_act, t_ln, dlogits, d_tln, d_tact, d_tpre, d_rows, logits, row_lse, row_argmax, row_aux};
  for (void* p : mlm) if (p) cudaFree(p);
  if (stats_host) cudaFreeHost(stats_host);
  if (comm && nccl) nccl->CommDestroy(comm);
  if (comm_stream) cudaStreamDestroy(comm_stream);
  if (ev_ready) cudaEventDestroy(ev_ready);
  if (ev_done) cudaEventDestroy(ev_done);
  if (ev_mid) cudaEventDestroy(ev_mid);
}

int Engine::ensure_mlm(int n) {
  if (n <= mlm_cap) return 0;
  MV_CUDA_CHECK(cudaDeviceSynchronize());
  void** ptrs[] = {&rows_h, &t_pre, &t_act, &t_ln, &dlogits, &d_tln, &d_tact, &d_tpre, &d_rows,
                   reinterpret_cast<void**>(&logits), reinterpret_cast<void**>(&row_lse), reinterpret_cast<void**>(&row_argmax),
                   reinterpret_cast<void**>(&row_aux)};
  for (void** p : ptrs) { if (*p) cudaFree(*p); *p = nullptr; }
  int cap = mlm_cap * 2 > n ? mlm_cap * 2 : n;
  cap = (cap + 127) / 128 * 128;
  const size_t H = cfg.hidden, c = cap;
  MV_CUDA_CHECK(cudaMalloc(&rows_h, c * H * es)); MV_CUDA_CHECK(cudaMalloc(&t_pre, c * H * es));
  MV_CUDA_CHECK(cudaMalloc(&t_act, c * H * es)); MV_CUDA_CHECK(cudaMalloc(&t_ln, c * H * es));
  MV_CUDA_CHECK(cudaMalloc(&dlogits, c * Vpad * es)); MV_CUDA_CHECK(cudaMalloc(&d_tln, c * H * es));
  MV_CUDA_CHECK(cudaMalloc(&d_tact, c * H * es)); MV_CUDA_CHECK(cudaMalloc(&d_tpre, c * H * es));
  MV_CUDA_CHECK(cudaMalloc(&d_rows, c * H * es));
  MV_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&logits), c * Vpad * sizeof(float)));
  MV_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&row_lse), c * sizeof(float)));
  MV_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&row_argmax), c * sizeof(int)));
  MV_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&row_aux), (2 * c + 4) * sizeof(float)));
  mlm_cap = cap;
  return 0;
}

// Y[M,N] = epi(X[M,K] . W[N,K]^T + b)
int Engine::linear_fwd(const void* X, int M, int K, int64_t w_off, int N, int64_t b_off, void* Y, int epi, void* pre,
                       const void* resid, int drop_on, uint32_t site, const DropoutCfg& dc, cudaStream_t s, int y_f32, long ldy,
                       int resid_f32) {
  GemmDesc d;
  d.M = M; d.N = N; d.K = K;
  d.A = X; d.lda = K; d.a_mn = 0;
  d.B = W(w_off); d.ldb = K; d.b_mn = 0;
  d.C = Y; d.ldc = ldy ? ldy : N; d.c_f32 = y_f32;
  d.C2 = pre; d.ldc2 = N;
  d.epi = epi; d.bias = b_off >= 0 ? params + b_off : nullptr;
  d.resid = resid; d.ldr = N; d.resid_f32 = resid_f32;
  d.drop_on = drop_on; d.drop_site = site; d.drop = dc;
  return gemm(d, s);
}

// dX[M,K] = epi(dY[M,N] . W[N,K])   (W read MN-major in place)
int Engine::linear_dgrad(const void* dY, long lddy, int M, int N, int64_t w_off, int K, void* dX, int epi, const void* extra,
                         cudaStream_t s) {
  GemmDesc d;
  d.M = M; d.N = K; d.K = N;
  d.A = dY; d.lda = lddy; d.a_mn = 0;
  d.B = W(w_off); d.ldb = K; d.b_mn = 1;
  d.C = dX; d.ldc = K;
  d.epi = epi; d.resid = extra; d.ldr = K; d.aux = extra; d.ldaux = K;
  return gemm(d, s);
}

// dW[N,K] += dY[M,N]^T . X[M,K]     (fp32, split-K reduce-add into the gradient arena)
int Engine::linear_wgrad(const void* dY, long lddy, const void* X, int M, int N, int K, int64_t w_off, cudaStream_t s) {
  GemmDesc d;
  d.M = N; d.N = K; d.K = M;
  d.A = dY; d.lda = lddy; d.a_mn = 1;
  d.B = X; d.ldb = K; d.b_mn = 1;
  d.C = grads + w_off; d.ldc = K; d.c_f32 = 1; d.accumulate = 1;
  d.epi = EPI_NONE;
  return gemm(d, s);
}

#define MV_TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

int Engine::forward(const mv_batch& b, cudaStream_t s) {
  MV_REQUIRE(params && grads && (f32 || shadow), "arenas not bound (mv_bind_arenas)");
  MV_REQUIRE(b.B > 0 && b.B <= cfg.max_batch, "batch %d exceeds capacity %d", b.B, cfg.max_batch);
  MV_REQUIRE(b.cls_tok && b.sep_tok && b.input_ids && b.segment && b.region_idx && b.mode && b.t_len && b.feats, "mv_batch: null input");
  MV_REQUIRE(b.n_lab >= 0 && b.n_lab <= b.B * T, "n_lab out of range");
  const int B = b.B, H = cfg.hidden, I = cfg.inter, N = cfg.num_image_embeds, M = B * L;
  const bool drop = b.train && cfg.dropout_p > 0.f;                    // hidden-state sites
  const bool drop_att = b.train && cfg.attn_dropout_p > 0.f, drop_emb = b.train && (cfg.dropout_p > 0.f || cfg.img_dropout_p > 0.f);
  const DropoutCfg dc = make_dropout(cfg.dropout_p, b.dropout_seed);
  const DropoutCfg dc_att = make_dropout(cfg.attn_dropout_p, b.dropout_seed), dc_img = make_dropout(cfg.img_dropout_p, b.dropout_seed);
  MV_TRY(ensure_mlm(b.n_lab));

  // --- visual tokens: sample regions, project 2048 -> H (models/image.py:57-69, cxrbert_origin.py:24) ---
  MV_TRY(gather_rows(b.feats, feats_g, b.region_idx, B * N, N, cfg.grid, cfg.img_hidden, f32, s));
  MV_TRY(linear_fwd(feats_g, B * N, cfg.img_hidden, lay.img_w, H, lay.img_b, proj, EPI_BIAS, nullptr, nullptr, 0, 0, dc, s));
  // --- joint embedding + LayerNorm + dropout, written straight into the concatenated [B, L, H] buffer ---
  EmbedArgs ea;
  ea.B = B; ea.L = L; ea.H = H; ea.N = N; ea.T = T; ea.A = A;
  ea.cls_tok = reinterpret_cast<const int64_t*>(b.cls_tok); ea.sep_tok = reinterpret_cast<const int64_t*>(b.sep_tok);
  ea.input_ids = reinterpret_cast<const int64_t*>(b.input_ids); ea.segment = reinterpret_cast<const int64_t*>(b.segment);
  ea.region_idx = reinterpret_cast<const int64_t*>(b.region_idx);
  ea.word = params + lay.word; ea.pos = params + lay.pos; ea.type = params + lay.type;
  ea.gamma = params + lay.emb_ln_g; ea.beta = params + lay.emb_ln_b; ea.eps = cfg.ln_eps;
  ea.V = cfg.vocab; ea.P = cfg.max_pos; ea.TV = cfg.type_vocab; ea.err = &stats->error_flags;
  ea.proj = proj; ea.emb_sum = emb_sum; ea.out = x[0]; ea.out32 = xres;
  ea.drop_on = drop_emb; ea.drop_site = SITE_EMB; ea.drop = dc; ea.drop_img = dc_img;
  MV_REQUIRE(b.sep_position >= 0 && b.sep_position < cfg.max_pos && b.prefix_type >= 0 && b.prefix_type < cfg.type_vocab,
             "mv_batch: sep_position %d / prefix_type %d out of range", b.sep_position, b.prefix_type);
  ea.sep_pos = b.sep_position; ea.prefix_type = b.prefix_type;
  MV_TRY(embed_ln_fwd(ea, f32, s));

  // --- encoder ---
  for (int l = 0; l < cfg.layers; ++l) {
    const int64_t base = lay.layer0 + l * lay.layer_stride;
    LayerWs& w = lw[l];
    MV_TRY(linear_fwd(x[l], M, H, base + lay.l_wqkv, 3 * H, base + lay.l_bqkv, w.qkv, EPI_BIAS, nullptr, nullptr, 0, 0, dc, s));
    AttnArgs aa;
    memset(&aa, 0, sizeof(aa));
    aa.B = B; aa.L = L; aa.nh = nh; aa.A = A; aa.mode = b.mode; aa.t_len = b.t_len; aa.qkv = w.qkv; aa.ctx = w.ctx; aa.lse = w.lse;
    aa.drop_on = drop_att; aa.drop_site = site_att(l); aa.drop = dc_att;
    MV_TRY(prof_begin(1, 4.0 * B * nh * static_cast<double>(L) * L * 64, s));
    MV_TRY(f32 ? attention_fwd_simt(aa, s) : attention_fwd_tc05(aa, s));
    MV_TRY(prof_end(s));
    // rf32: y1 / y2 and the residual operands are fp32 (the epilogue adds the unrounded LayerNorm output and stores the
    // unrounded sum), so the residual stream sees no bf16 rounding between layers — only GEMM operands are bf16
    MV_TRY(linear_fwd(w.ctx, M, H, base + lay.l_wo, H, base + lay.l_bo, w.y1, EPI_BIAS_RESID, nullptr, rf32 ? static_cast<const void*>(xres) : x[l],
                      drop, site_h1(l), dc, s, rf32, 0, rf32));
    MV_TRY(ln_fwd(w.y1, w.x1, params + base + lay.l_ln1_g, params + base + lay.l_ln1_b, M, H, cfg.ln_eps, 0, 0, dc, f32, s, rf32, x1res));
    // h1 receives gelu'(pre-activation) — what backward needs — instead of the pre-activation itself
    MV_TRY(linear_fwd(w.x1, M, H, base + lay.l_w1, I, base + lay.l_b1, w.g1, EPI_BIAS_GELU_GRAD, w.h1, nullptr, 0, 0, dc, s));
    MV_TRY(linear_fwd(w.g1, M, I, base + lay.l_w2, H, base + lay.l_b2, w.y2, EPI_BIAS_RESID, nullptr, rf32 ? static_cast<const void*>(x1res) : w.x1,
                      drop, site_h2(l), dc, s, rf32, 0, rf32));
    MV_TRY(ln_fwd(w.y2, x[l + 1], params + base + lay.l_ln2_g, params + base + lay.l_ln2_b, M, H, cfg.ln_eps, 0, 0, dc, f32, s, rf32, xres));
  }
  const void* seq = x[cfg.layers];

  // --- pooler + ITM head + CE (cxrbert_origin.py:130,164-173; train_origin.py:63,123,133-136) ---
  MV_TRY(gather_rows(seq, cls_rows, zero_idx, B, 1, L, H, f32, s));
  MV_TRY(linear_fwd(cls_rows, B, H, lay.pool_w, H, lay.pool_b, pooled, EPI_BIAS_TANH, nullptr, nullptr, 0, 0, dc, s));
  ItmArgs ia;
  ia.B = B; ia.H = H; ia.pooled = pooled; ia.w = params + lay.itm_w; ia.b = params + lay.itm_b;
  ia.labels = reinterpret_cast<const int64_t*>(b.is_aligned); ia.gscale = b.inv_batch_global; ia.logits = itm_logits;
  ia.count_dev = b.global_counts ? b.global_counts + 1 : nullptr;
  ia.loss_sum = &stats->itm_loss_sum; ia.correct = &stats->itm_correct;
  ia.d_pre = (b.train && b.is_aligned) ? d_pre : nullptr; ia.dw = grads + lay.itm_w; ia.db = grads + lay.itm_b;
  MV_TRY(itm_head_fwd_bwd(ia, f32, s));

  // --- MLM head on the labelled rows only (cxrbert_origin.py:205-248; train_origin.py:62,120,138-146) ---
  if (b.n_lab > 0) {
    MV_REQUIRE(b.lab_rows && b.lab_labels, "mv_batch: n_lab > 0 needs lab_rows / lab_labels");
    const int n = b.n_lab;
    MV_TRY(mlm_head_rows(reinterpret_cast<const int64_t*>(b.lab_rows), n, dc, s));
    MV_TRY(linear_fwd(t_ln, n, H, lay.word, cfg.vocab, lay.mlm_bias, logits, EPI_BIAS, nullptr, nullptr, 0, 0, dc, s, 1, Vpad));
    CeArgs ca;
    ca.n = n; ca.V = cfg.vocab; ca.ldv = Vpad; ca.logits = logits; ca.labels = reinterpret_cast<const int64_t*>(b.lab_labels);
    ca.correct = &stats->mlm_correct; ca.row_lse = row_lse; ca.row_argmax = row_argmax; ca.err = &stats->error_flags;
    if (b.drop_worst_keep > 0) {
      // model.py:1003-1010: which samples count is only known once every row's loss is; so one loss-only CE pass, the
      // selection, then the gradient pass with row_scale = kept * weight / (kept weights + 1e-5)
      float* row_loss = row_aux; float* row_scale = row_aux + mlm_cap; float* scratch = row_aux + 2 * static_cast<size_t>(mlm_cap);
      ca.dlogits = nullptr; ca.loss_sum = scratch; ca.row_weight = nullptr; ca.row_loss = row_loss;
      MV_TRY(mlm_ce_fwd_bwd(ca, f32, s));
      DropWorstArgs dw;
      dw.n = n; dw.B = B; dw.L = L; dw.keep = b.drop_worst_keep;
      dw.row_loss = row_loss; dw.row_weight = b.lab_weights; dw.rows = reinterpret_cast<const int64_t*>(b.lab_rows);
      dw.row_scale = row_scale; dw.loss_sum = &stats->mlm_loss_sum;
      MV_TRY(drop_worst_select(dw, s));
      if (b.train) {
        ca.dlogits = dlogits; ca.gscale = b.inv_n_lab_global; ca.count_dev = nullptr; ca.row_weight = row_scale; ca.row_loss = nullptr;
        ca.correct = reinterpret_cast<int*>(scratch + 1); ca.row_lse = nullptr; ca.row_argmax = nullptr;
        MV_TRY(mlm_ce_fwd_bwd(ca, f32, s));
      }
    } else {
      ca.dlogits = b.train ? dlogits : nullptr; ca.gscale = b.inv_n_lab_global;
      ca.count_dev = b.global_counts;
      ca.loss_sum = &stats->mlm_loss_sum;
      ca.row_weight = b.lab_weights;
      MV_TRY(mlm_ce_fwd_bwd(ca, f32, s));
    }
  }
  return 0;
}

// MLM head up to the decoder input for the given rows of the final hidden states (cxrbert_origin.py:205-218)
int Engine::mlm_head_rows(const int64_t* rows, int n, const DropoutCfg& dc, cudaStream_t s) {
  const int H = cfg.hidden;
  MV_TRY(ensure_mlm(n));
  MV_TRY(gather_rows(x[cfg.layers], rows_h, rows, n, n, 0, H, f32, s));
  MV_TRY(linear_fwd(rows_h, n, H, lay.mlm_tw, H, lay.mlm_tb, t_act, EPI_BIAS_GELU_GRAD, t_pre, nullptr, 0, 0, dc, s));   // t_pre = gelu'
  MV_TRY(ln_fwd(t_act, t_ln, params + lay.mlm_ln_g, params + lay.mlm_ln_b, n, H, cfg.head_ln_eps, 0, 0, dc, f32, s));
  return 0;
}

// Backward from gradients the CALLER computed (the torch.autograd drop-in path: loss functions run in PyTorch on the
// [B, L, V] / [B, 2] outputs, models/train_origin.py:118-130).  Only rows of the logit gradient that are not identically
// zero are handed over (CrossEntropy(ignore_index=-100) zeroes every unlabelled position); the head is re-run for those rows
// to rebuild what its backward needs, then the common backward runs.
int Engine::backward_external(const mv_batch& b_in, const mv_external_grads& g, int allreduce, cudaStream_t s) {
  MV_REQUIRE(b_in.train, "mv_backward_external needs a batch forwarded with train=1");
  MV_REQUIRE(g.n_rows >= 0 && g.n_rows <= b_in.B * L, "mv_backward_external: n_rows out of range");
  MV_REQUIRE(g.n_rows == 0 || (g.rows && g.dlogits), "mv_backward_external: rows / dlogits missing");
  const int H = cfg.hidden, B = b_in.B;
  const DropoutCfg dc0 = make_dropout(0.f, 0);
  mv_batch b = b_in;
  b.n_lab = g.n_rows; b.lab_rows = g.rows; b.lab_labels = nullptr;
  if (g.n_rows > 0) MV_TRY(mlm_head_rows(reinterpret_cast<const int64_t*>(g.rows), g.n_rows, dc0, s));
  // pooler pre-activation gradient: from the ITM logits' gradient (through the ITM linear layer) and / or from a gradient
  // on the pooled output itself
  bool have_pre = false;
  if (g.d_itm) {
    ItmArgs ia;
    ia.B = B; ia.H = H; ia.pooled = pooled; ia.w = params + lay.itm_w; ia.b = params + lay.itm_b; ia.labels = nullptr;
    ia.gscale = 1.f; ia.logits = itm_logits; ia.loss_sum = &stats->itm_loss_sum; ia.correct = &stats->itm_correct;
    ia.d_pre = d_pre; ia.dw = grads + lay.itm_w; ia.db = grads + lay.itm_b; ia.ext_dlogits = g.d_itm;
    MV_TRY(itm_head_fwd_bwd(ia, f32, s));
    have_pre = true;
  }
  if (g.d_pooled) {
    MV_TRY(tanh_bwd(g.d_pooled, pooled, d_pre, static_cast<long>(B) * H, have_pre ? 1 : 0, f32, s));
    have_pre = true;
  }
  mv_batch bb = b;
  static const int64_t kDummy = 0;
  bb.is_aligned = have_pre ? &kDummy : nullptr;      // backward() only tests it for null (d_pre is ready)
  ext_dlogits = g.n_rows > 0 ? g.dlogits : nullptr;
  ext_dseq = g.d_seq;
  const int rc = backward(bb, allreduce, s);
  ext_dlogits = nullptr;
  ext_dseq = nullptr;
  return rc;
}

int Engine::bucket_done(size_t idx, int allreduce, cudaStream_t s) {
  if (!allreduce || world <= 1) return 0;
  MV_REQUIRE(comm != nullptr, "allreduce requested but mv_comm_init was not called");
  const Bucket& bk = buckets[idx];
  MV_CUDA_CHECK(cudaEventRecord(ev_ready, s));
  MV_CUDA_CHECK(cudaStreamWaitEvent(comm_stream, ev_ready, 0));
  MV_TRY(engine_allreduce_bucket(this, grads + bk.offset, bk.count, comm_stream));
  comm_pending = true;
  if (idx + 2 == buckets.size()) {          // everything except the last bucket (embeddings) is now queued on comm_stream
    MV_CUDA_CHECK(cudaEventRecord(ev_mid, comm_stream));
    mid_recorded = true;
  }
  set_reserved_sms(comm_ctas(world));       // GEMMs launched from now on leave room for the collective's CTAs
  return 0;
}

int Engine::backward(const mv_batch& b, int allreduce, cudaStream_t s) {
  MV_REQUIRE(b.train, "mv_backward needs a batch forwarded with train=1");
  const int B = b.B, H = cfg.hidden, I = cfg.inter, N = cfg.num_image_embeds, M = B * L;
  const bool drop = cfg.dropout_p > 0.f, drop_att = cfg.attn_dropout_p > 0.f, drop_emb = cfg.dropout_p > 0.f || cfg.img_dropout_p > 0.f;
  const DropoutCfg dc = make_dropout(cfg.dropout_p, b.dropout_seed);
  const DropoutCfg dc_att = make_dropout(cfg.attn_dropout_p, b.dropout_seed);
  void* P = dxa; void* Q = dxb; void* R = dxc;
  if (ext_dseq) MV_CUDA_CHECK(cudaMemcpyAsync(P, ext_dseq, static_cast<size_t>(M) * H * es, cudaMemcpyDeviceToDevice, s));
  else MV_CUDA_CHECK(cudaMemsetAsync(P, 0, static_cast<size_t>(M) * H * es, s));   // d(seq): only labelled + [CLS] rows are non-zero
  const void* dlogits = ext_dlogits ? ext_dlogits : this->dlogits;                // shadows the member on purpose

  // --- MLM head backward ---
  if (b.n_lab > 0) {
    const int n = b.n_lab;
    MV_TRY(colsum_add(dlogits, Vpad, n, Vpad, grads + lay.mlm_bias, f32, s));
    {  // tied decoder weight: dE[V,H] += dlogits^T . t_ln   (the embedding-lookup part is added at the end)
      GemmDesc d;
      d.M = cfg.vocab; d.N = H; d.K = n;
      d.A = dlogits; d.lda = Vpad; d.a_mn = 1;
      d.B = t_ln; d.ldb = H; d.b_mn = 1;
      d.C = grads + lay.word; d.ldc = H; d.c_f32 = 1; d.accumulate = 1;
      MV_TRY(gemm(d, s));
    }
    MV_TRY(linear_dgrad(dlogits, Vpad, n, cfg.vocab, lay.word, H, d_tln, EPI_NONE, nullptr, s));
    MV_TRY(ln_bwd(d_tln, t_act, params + lay.mlm_ln_g, d_tact, nullptr, grads + lay.mlm_ln_g, grads + lay.mlm_ln_b, nullptr, n, H,
                  cfg.head_ln_eps, 0, 0, 0, dc, f32, s));
    MV_TRY(mul_elem(d_tact, t_pre, d_tpre, static_cast<long>(n) * H, f32, s));
    MV_TRY(colsum_add(d_tpre, H, n, H, grads + lay.mlm_tb, f32, s));
    MV_TRY(linear_wgrad(d_tpre, H, rows_h, n, H, H, lay.mlm_tw, s));
    MV_TRY(linear_dgrad(d_tpre, H, n, H, lay.mlm_tw, H, d_rows, EPI_NONE, nullptr, s));
    MV_TRY(scatter_rows(d_rows, P, reinterpret_cast<const int64_t*>(b.lab_rows), n, n, 0, H, ext_dseq ? 1 : 0, f32, s));
  }
  // --- pooler backward (d_pre was produced by the ITM kernel in forward) ---
  if (b.is_aligned) {
    MV_TRY(colsum_add(d_pre, H, B, H, grads + lay.pool_b, f32, s));
    MV_TRY(linear_wgrad(d_pre, H, cls_rows, B, H, H, lay.pool_w, s));
    MV_TRY(linear_dgrad(d_pre, H, B, H, lay.pool_w, H, d_cls, EPI_NONE, nullptr, s));
    MV_TRY(scatter_rows(d_cls, P, zero_idx, B, 1, L, H, 1, f32, s));
  }
  MV_TRY(bucket_done(0, allreduce, s));

  // --- encoder layers, last to first ---
  for (int l = cfg.layers - 1; l >= 0; --l) {
    const int64_t base = lay.layer0 + l * lay.layer_stride;
    LayerWs& w = lw[l];
    float* g = grads + base;
    // P = d(layer output).  LN2: y2 -> x[l+1]
    void* dY2d = drop ? R : Q;
    MV_TRY(ln_bwd(P, w.y2, params + base + lay.l_ln2_g, Q, drop ? R : nullptr, g + lay.l_ln2_g, g + lay.l_ln2_b, g + lay.l_b2, M, H,
                  cfg.ln_eps, 0, drop, site_h2(l), dc, f32, s, rf32));
    MV_TRY(linear_wgrad(dY2d, H, w.g1, M, H, I, base + lay.l_w2, s));
    MV_TRY(linear_dgrad(dY2d, H, M, H, base + lay.l_w2, I, dh1, EPI_MUL, w.h1, s));
    MV_TRY(colsum_add(dh1, I, M, I, g + lay.l_b1, f32, s));
    MV_TRY(linear_wgrad(dh1, I, w.x1, M, I, H, base + lay.l_w1, s));
    MV_TRY(linear_dgrad(dh1, I, M, I, base + lay.l_w1, H, P, EPI_RESID, Q, s));          // P = d(x1)
    // LN1: y1 -> x1
    void* dY1d = drop ? R : Q;
    MV_TRY(ln_bwd(P, w.y1, params + base + lay.l_ln1_g, Q, drop ? R : nullptr, g + lay.l_ln1_g, g + lay.l_ln1_b, g + lay.l_bo, M, H,
                  cfg.ln_eps, 0, drop, site_h1(l), dc, f32, s, rf32));
    MV_TRY(linear_wgrad(dY1d, H, w.ctx, M, H, H, base + lay.l_wo, s));
    MV_TRY(linear_dgrad(dY1d, H, M, H, base + lay.l_wo, H, dctx, EPI_NONE, nullptr, s));
    AttnArgs aa;
    memset(&aa, 0, sizeof(aa));
    aa.B = B; aa.L = L; aa.nh = nh; aa.A = A; aa.mode = b.mode; aa.t_len = b.t_len; aa.qkv = w.qkv; aa.ctx = w.ctx; aa.lse = w.lse;
    aa.dctx = dctx; aa.dqkv = dqkv; aa.dq_acc = dq_acc; aa.delta = delta; aa.dq_part = dq_part;
    aa.drop_on = drop_att; aa.drop_site = site_att(l); aa.drop = dc_att;
    MV_TRY(prof_begin(2, 10.0 * B * nh * static_cast<double>(L) * L * 64, s));
    MV_TRY(f32 ? attention_bwd_simt(aa, s) : attention_bwd_tc05(aa, s));
    MV_TRY(prof_end(s));
    MV_TRY(colsum_add(dqkv, 3 * H, M, 3 * H, g + lay.l_bqkv, f32, s));
    MV_TRY(linear_wgrad(dqkv, 3 * H, x[l], M, 3 * H, H, base + lay.l_wqkv, s));
    MV_TRY(linear_dgrad(dqkv, 3 * H, M, 3 * H, base + lay.l_wqkv, H, P, EPI_RESID, Q, s));   // P = d(x[l])
    MV_TRY(bucket_done(static_cast<size_t>(cfg.layers - l), allreduce, s));
  }

  // --- embeddings: dropout-bwd + LN-bwd, scatter into the tables, image-projection wgrad ---
  LnAltDrop alt;
  alt.period = L; alt.lo = 1; alt.hi = N + 1; alt.drop = make_dropout(cfg.img_dropout_p, b.dropout_seed);
  MV_TRY(ln_bwd(P, emb_sum, params + lay.emb_ln_g, Q, nullptr, grads + lay.emb_ln_g, grads + lay.emb_ln_b, nullptr, M, H, cfg.ln_eps,
                drop_emb, 0, SITE_EMB, dc, f32, s, 0, &alt));
  EmbedBwdArgs eb;
  eb.B = B; eb.L = L; eb.H = H; eb.N = N; eb.T = T; eb.A = A; eb.V = cfg.vocab;
  eb.cls_tok = reinterpret_cast<const int64_t*>(b.cls_tok); eb.sep_tok = reinterpret_cast<const int64_t*>(b.sep_tok);
  eb.input_ids = reinterpret_cast<const int64_t*>(b.input_ids); eb.segment = reinterpret_cast<const int64_t*>(b.segment);
  eb.region_idx = reinterpret_cast<const int64_t*>(b.region_idx);
  eb.dsum = Q; eb.d_word = grads + lay.word; eb.d_pos = grads + lay.pos; eb.d_type = grads + lay.type; eb.d_proj = dproj; eb.pad_id = b.pad_lookup_grad ? -1 : 0;
  eb.sep_pos = b.sep_position; eb.prefix_type = b.prefix_type; eb.TV = cfg.type_vocab; eb.P = cfg.max_pos;
  MV_TRY(embed_bwd_scatter(eb, f32, s));
  MV_TRY(colsum_add(dproj, H, B * N, H, grads + lay.img_b, f32, s));
  MV_TRY(linear_wgrad(dproj, H, feats_g, B * N, H, cfg.img_hidden, lay.img_w, s));
  MV_TRY(bucket_done(buckets.size() - 1, allreduce, s));
  return 0;
}

// Parameter tensors the report-generation fine-tune updates (finetune.py:383-395): everything with a gradient, i.e. all
// but the pooler and the ITM head (BertAdam skips p.grad is None, optimization.py:128-129); q/k/v are separate
// Parameters in the reference, so each gets its own clipping norm.  decay = name has no 'bias' / 'LayerNorm' part.
int Engine::bert_adam(float lr, float beta1, float beta2, float eps, float weight_decay, float max_grad_norm, cudaStream_t s) {
  MV_REQUIRE(params && grads && adam_m && adam_v, "mv_bert_adam_step: arenas (incl. Adam moments) not bound");
  if (!adam_chunks) {
    struct T { int64_t off, n; int decay; };
    std::vector<T> ts;
    const int64_t H = cfg.hidden, I = cfg.inter, V = cfg.vocab;
    ts.push_back({lay.word, V * H, 1}); ts.push_back({lay.pos, static_cast<int64_t>(cfg.max_pos) * H, 1});
    ts.push_back({lay.type, static_cast<int64_t>(cfg.type_vocab) * H, 1});
    ts.push_back({lay.emb_ln_g, H, 0}); ts.push_back({lay.emb_ln_b, H, 0});
    ts.push_back({lay.img_w, H * cfg.img_hidden, 1}); ts.push_back({lay.img_b, H, 0});
    for (int l = 0; l < cfg.layers; ++l) {
      const int64_t b = lay.layer0 + l * lay.layer_stride;
      for (int j = 0; j < 3; ++j) ts.push_back({b + lay.l_wqkv + j * H * H, H * H, 1});
      for (int j = 0; j < 3; ++j) ts.push_back({b + lay.l_bqkv + j * H, H, 0});
      ts.push_back({b + lay.l_wo, H * H, 1}); ts.push_back({b + lay.l_bo, H, 0});
      ts.push_back({b + lay.l_ln1_g, H, 0}); ts.push_back({b + lay.l_ln1_b, H, 0});
      ts.push_back({b + lay.l_w1, I * H, 1}); ts.push_back({b + lay.l_b1, I, 0});
      ts.push_back({b + lay.l_w2, H * I, 1}); ts.push_back({b + lay.l_b2, H, 0});
      ts.push_back({b + lay.l_ln2_g, H, 0}); ts.push_back({b + lay.l_ln2_b, H, 0});
    }
    ts.push_back({lay.mlm_bias, V, 0}); ts.push_back({lay.mlm_tw, H * H, 1}); ts.push_back({lay.mlm_tb, H, 0});
    ts.push_back({lay.mlm_ln_g, H, 0}); ts.push_back({lay.mlm_ln_b, H, 0});
    std::vector<AdamChunk> ch;
    const int64_t kChunk = 32768;
    for (size_t t = 0; t < ts.size(); ++t)
      for (int64_t o = 0; o < ts[t].n; o += kChunk)
        ch.push_back({static_cast<long>(ts[t].off + o), static_cast<int>(ts[t].n - o < kChunk ? ts[t].n - o : kChunk), static_cast<int>(t), ts[t].decay});
    void* p = nullptr;
    MV_TRY(alloc(&p, ch.size() * sizeof(AdamChunk)));
    adam_chunks = static_cast<AdamChunk*>(p);
    MV_TRY(alloc(&p, ts.size() * sizeof(float)));
    adam_sumsq = static_cast<float*>(p);
    MV_CUDA_CHECK(cudaMemcpy(adam_chunks, ch.data(), ch.size() * sizeof(AdamChunk), cudaMemcpyHostToDevice));
    n_adam_chunks = static_cast<int>(ch.size());
    n_adam_tensors = static_cast<int>(ts.size());
  }
  BertAdamArgs a;
  a.p = params; a.g = grads; a.m = adam_m; a.v = adam_v; a.shadow = shadow;
  a.chunks = adam_chunks; a.n_chunks = n_adam_chunks; a.sumsq = adam_sumsq; a.n_tensors = n_adam_tensors;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.max_grad_norm = max_grad_norm;
  return bert_adam_step(a, s);
}

// prediction_scores for EVERY position (reference semantics, models/cxrbert_origin.py:147): [B*L, ld] fp32
int Engine::full_logits(const mv_batch& b, float* out, int64_t ld, cudaStream_t s) {
  const int H = cfg.hidden, M = b.B * L;
  MV_REQUIRE(ld >= cfg.vocab && ld % 4 == 0, "full_logits: ld must be >= vocab and a multiple of 4");
  // in row chunks through the head's existing [mlm_cap, H] buffers (growing them to B*L rows would also grow the
  // [rows, V] logit / gradient buffers of the training path: ~5 GB at B = 64)
  const DropoutCfg dc = make_dropout(0.f, 0);
  const char* seq = static_cast<const char*>(x[cfg.layers]);
  for (int r0 = 0; r0 < M; r0 += mlm_cap) {
    const int n = M - r0 < mlm_cap ? M - r0 : mlm_cap;
    MV_TRY(linear_fwd(seq + static_cast<size_t>(r0) * H * es, n, H, lay.mlm_tw, H, lay.mlm_tb, t_act, EPI_BIAS_GELU, nullptr, nullptr, 0, 0, dc, s));
    MV_TRY(ln_fwd(t_act, t_ln, params + lay.mlm_ln_g, params + lay.mlm_ln_b, n, H, cfg.head_ln_eps, 0, 0, dc, f32, s));
    MV_TRY(linear_fwd(t_ln, n, H, lay.word, cfg.vocab, lay.mlm_bias, out + static_cast<size_t>(r0) * ld, EPI_BIAS, nullptr, nullptr, 0, 0, dc, s, 1, ld));
  }
  return 0;
}

}  // namespace mv
