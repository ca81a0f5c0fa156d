// gemm_test.cu — standalone GPU check of the tcgen05 GEMM (all operand-major combos, epilogues, tails, split-K)
// and the fp32 SIMT GEMM against a double-precision CPU contraction of the same (bf16-rounded) inputs.
// Build: make -C csrc gemm_test ; run on a B200: ./gemm_test [--perf]
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../gemm.h"

using namespace mv;

static uint32_t g_seed = 12345;
static float frand() {
  g_seed = g_seed * 1664525u + 1013904223u;
  return ((g_seed >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}
static float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

struct Case {
  const char* name;
  int M, N, K, a_mn, b_mn, c_f32, accumulate, epi, use_c2, drop;
};

static double gelu_d(double x) { return 0.5 * x * (1.0 + erf(x / sqrt(2.0))); }
static double gelu_grad_d(double x) {
  return 0.5 * (1.0 + erf(x / sqrt(2.0))) + x * exp(-0.5 * x * x) / sqrt(2.0 * M_PI);
}

static int g_pair = -1, g_bn = 0;   // forced CTA-pair mode / N tile of the tcgen05 path (-1 / 0 = automatic)

static int run_case(const Case& c, bool fp32_mode, int n_samples) {
  const int M = c.M, N = c.N, K = c.K;
  // logical A[m][k], B[n][k]
  std::vector<float> A((size_t)M * K), B((size_t)N * K), bias(N), R((size_t)M * N), Cinit((size_t)M * N);
  for (auto& v : A) v = bf16_round(frand());
  for (auto& v : B) v = bf16_round(frand());
  for (auto& v : bias) v = frand();
  for (auto& v : R) v = bf16_round(frand() * 2.f);
  for (auto& v : Cinit) v = frand();
  const long lda = c.a_mn ? M : K, ldb = c.b_mn ? N : K;
  std::vector<float> Ast((size_t)M * K), Bst((size_t)N * K);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) Ast[c.a_mn ? (size_t)k * M + m : (size_t)m * K + k] = A[(size_t)m * K + k];
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) Bst[c.b_mn ? (size_t)k * N + n : (size_t)n * K + k] = B[(size_t)n * K + k];

  const size_t esz = fp32_mode ? 4 : 2;
  auto upload = [&](const std::vector<float>& h, bool as_act) -> void* {
    void* d = nullptr;
    if (as_act && !fp32_mode) {
      std::vector<bf16> hb(h.size());
      for (size_t i = 0; i < h.size(); ++i) hb[i] = __float2bfloat16_rn(h[i]);
      cudaMalloc(&d, hb.size() * 2);
      cudaMemcpy(d, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    } else {
      cudaMalloc(&d, h.size() * 4);
      cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    }
    return d;
  };
  void* dA = upload(Ast, true);
  void* dB = upload(Bst, true);
  void* dR = upload(R, true);
  float* dbias = (float*)upload(bias, false);
  void* dC = nullptr;
  void* dC2 = nullptr;
  const size_t csz = (size_t)M * N * (c.c_f32 ? 4 : esz);
  cudaMalloc(&dC, csz);
  if (c.c_f32) cudaMemcpy(dC, Cinit.data(), csz, cudaMemcpyHostToDevice); else cudaMemset(dC, 0xFF, csz);
  if (c.use_c2) { cudaMalloc(&dC2, (size_t)M * N * esz); cudaMemset(dC2, 0xFF, (size_t)M * N * esz); }

  GemmDesc d;
  d.M = M; d.N = N; d.K = K;
  d.A = dA; d.lda = lda; d.a_mn = c.a_mn;
  d.B = dB; d.ldb = ldb; d.b_mn = c.b_mn;
  d.C = dC; d.ldc = N; d.c_f32 = c.c_f32; d.accumulate = c.accumulate;
  d.C2 = dC2; d.ldc2 = N;
  d.epi = c.epi; d.bias = dbias; d.resid = dR; d.ldr = N; d.aux = dR; d.ldaux = N;
  d.drop_on = c.drop; d.drop_site = 7; d.drop = make_dropout(0.1f, 99);
  d.pair = g_pair; d.bn = g_bn;
  int rc = fp32_mode ? gemm_f32_simt(d, 0) : gemm_bf16_tc05(d, 0);
  if (rc) { printf("  %-34s launch error: %s\n", c.name, last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  %-34s CUDA error: %s\n", c.name, cudaGetErrorString(e)); exit(3); }

  std::vector<float> out((size_t)M * N), out2;
  auto download = [&](void* dptr, std::vector<float>& h, bool f32) {
    h.resize((size_t)M * N);
    if (f32) { cudaMemcpy(h.data(), dptr, h.size() * 4, cudaMemcpyDeviceToHost); }
    else {
      std::vector<bf16> hb(h.size());
      cudaMemcpy(hb.data(), dptr, hb.size() * 2, cudaMemcpyDeviceToHost);
      for (size_t i = 0; i < h.size(); ++i) h[i] = __bfloat162float(hb[i]);
    }
  };
  download(dC, out, c.c_f32 || fp32_mode);
  if (c.use_c2) download(dC2, out2, fp32_mode);

  double max_err = 0, max_err2 = 0;
  long dropped = 0, checked = 0;
  const bool full = (long)M * N <= n_samples;
  const long total = full ? (long)M * N : n_samples;
  for (long t = 0; t < total; ++t) {
    long idx = full ? t : (long)((double)(frand() + 0.5) * ((double)M * N - 1));
    // always include the far corner
    if (!full && t == 0) idx = (long)M * N - 1;
    const int m = idx / N, n = idx % N;
    double acc = 0;
    for (int k = 0; k < K; ++k) acc += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
    double pre = acc, v = acc;
    const int epi = c.epi;
    if (epi == EPI_BIAS || epi == EPI_BIAS_GELU || epi == EPI_BIAS_RESID || epi == EPI_BIAS_TANH || epi == EPI_BIAS_GELU_GRAD) v += bias[n];
    pre = v;
    if (epi == EPI_BIAS_GELU) v = gelu_d(v);
    else if (epi == EPI_BIAS_GELU_GRAD) { pre = gelu_grad_d(v); v = gelu_d(v); }   // C2 holds the derivative
    else if (epi == EPI_MUL) v *= R[idx];
    else if (epi == EPI_BIAS_TANH) v = tanh(v);
    else if (epi == EPI_BIAS_RESID || epi == EPI_RESID) {
      if (c.drop) {
        // dropout: accept either the kept (scaled) or the dropped value, count drops
        const double kept = v * (double)d.drop.scale + R[idx], drp = R[idx];
        const double got = out[idx];
        if (fabs(got - drp) < fabs(got - kept)) { v = drp; ++dropped; } else v = kept;
        ++checked;
      } else v += R[idx];
    } else if (epi == EPI_DGELU) v *= gelu_grad_d(R[idx]);
    if (c.accumulate) v += Cinit[idx];
    const double tol_scale = fmax(1.0, fabs(v));
    max_err = fmax(max_err, fabs(out[idx] - v) / tol_scale);
    if (c.use_c2) max_err2 = fmax(max_err2, fabs(out2[idx] - pre) / fmax(1.0, fabs(pre)));
  }
  const double tol = (fp32_mode || c.c_f32) ? 2e-4 : 1.2e-2;
  bool ok = max_err < tol && (!c.use_c2 || max_err2 < 1.2e-2);
  if (c.drop) {
    const double rate = (double)dropped / (double)(checked ? checked : 1);
    ok = ok && fabs(rate - 0.1) < 0.02;
    printf("  %-34s [%s] max_rel_err=%.3e drop_rate=%.4f %s\n", c.name, fp32_mode ? "f32" : "tc05", max_err, rate,
           ok ? "PASS" : "FAIL");
  } else {
    printf("  %-34s [%s] max_rel_err=%.3e%s %s\n", c.name, fp32_mode ? "f32" : "tc05", max_err,
           c.use_c2 ? " (+C2)" : "", ok ? "PASS" : "FAIL");
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dR); cudaFree(dbias); cudaFree(dC); if (dC2) cudaFree(dC2);
  return ok ? 0 : 1;
}

static void perf(const char* name, int M, int N, int K, int a_mn, int b_mn, int c_f32, int acc, int epi) {
  void *dA, *dB, *dC, *dR; float* dbias;
  cudaMalloc(&dA, (size_t)M * K * 2); cudaMalloc(&dB, (size_t)N * K * 2);
  cudaMalloc(&dC, (size_t)M * N * 4); cudaMalloc(&dR, (size_t)M * N * 2); cudaMalloc(&dbias, N * 4);
  {   // random operands: all-zero inputs draw far less power and flatter the clocks
    std::vector<bf16> h((size_t)std::max(std::max(M, N) * (size_t)K, (size_t)M * N));
    for (auto& v : h) v = __float2bfloat16_rn(frand());
    cudaMemcpy(dA, h.data(), (size_t)M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, h.data(), (size_t)N * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dR, h.data(), (size_t)M * N * 2, cudaMemcpyHostToDevice);
  }
  cudaMemset(dbias, 0, N * 4); cudaMemset(dC, 0, (size_t)M * N * 4);
  GemmDesc d;
  d.M = M; d.N = N; d.K = K; d.A = dA; d.lda = a_mn ? M : K; d.a_mn = a_mn; d.B = dB; d.ldb = b_mn ? N : K; d.b_mn = b_mn;
  d.C = dC; d.ldc = N; d.c_f32 = c_f32; d.accumulate = acc; d.epi = epi; d.bias = dbias; d.resid = dR; d.ldr = N;
  d.aux = dR; d.ldaux = N; d.pair = g_pair; d.bn = g_bn;
  void* dC2 = nullptr;
  if (epi == EPI_BIAS_GELU || epi == EPI_BIAS_GELU_GRAD) { cudaMalloc(&dC2, (size_t)M * N * 2); d.C2 = dC2; d.ldc2 = N; }   // as the engine runs it
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) gemm_bf16_tc05(d, 0);
  cudaEventRecord(e0);
  const int iters = 20;
  for (int i = 0; i < iters; ++i) gemm_bf16_tc05(d, 0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
  printf("  perf %-28s M=%d N=%d K=%d: %.3f ms  %.1f TFLOP/s\n", name, M, N, K, ms, 2.0 * M * N * K / ms / 1e9);
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dR); cudaFree(dbias); if (dC2) cudaFree(dC2);
}

static void perf_one(int which) {
  switch (which) {
    case 0: perf("QKV fwd", 27904, 2304, 768, 0, 0, 0, 0, EPI_BIAS); break;
    case 1: perf("out-proj fwd", 27904, 768, 768, 0, 0, 0, 0, EPI_BIAS_RESID); break;
    case 2: perf("FFN1 fwd", 27904, 3072, 768, 0, 0, 0, 0, EPI_BIAS_GELU); break;
    case 3: perf("FFN2 fwd", 27904, 768, 3072, 0, 0, 0, 0, EPI_BIAS_RESID); break;
    case 4: perf("FFN2 dgrad", 27904, 3072, 768, 0, 1, 0, 0, EPI_DGELU); break;
    case 5: perf("FFN1 dgrad", 27904, 768, 3072, 0, 1, 0, 0, EPI_RESID); break;
    case 6: perf("FFN1 wgrad", 3072, 768, 27904, 1, 1, 1, 1, EPI_NONE); break;
    case 7: perf("QKV wgrad", 2304, 768, 27904, 1, 1, 1, 1, EPI_NONE); break;
    case 8: perf("FFN1 fwd bias only", 27904, 3072, 768, 0, 0, 0, 0, EPI_BIAS); break;
    case 9: perf("FFN1 fwd gelu+gelu'", 27904, 3072, 768, 0, 0, 0, 0, EPI_BIAS_GELU_GRAD); break;
    case 10: perf("FFN2 dgrad mul", 27904, 3072, 768, 0, 1, 0, 0, EPI_MUL); break;
    default: perf("out wgrad", 768, 768, 27904, 1, 1, 1, 1, EPI_NONE); break;
  }
}

int main(int argc, char** argv) {
  if (argc > 3 && !strcmp(argv[1], "--perf-one")) g_pair = atoi(argv[3]);
  if (argc > 2 && !strcmp(argv[1], "--perf-one")) { perf_one(atoi(argv[2])); return 0; }
  bool do_perf = argc > 1 && (!strcmp(argv[1], "--perf") || !strcmp(argv[1], "--perf-only"));
  const bool perf_only = argc > 1 && !strcmp(argv[1], "--perf-only");
  int fails = 0;
  const Case cases[] = {
      {"TN 128x192x64 none", 128, 192, 64, 0, 0, 0, 0, EPI_NONE, 0, 0},
      {"TN 128x192x256 none", 128, 192, 256, 0, 0, 0, 0, EPI_NONE, 0, 0},
      {"TN 300x200x136 bias (tails)", 300, 200, 136, 0, 0, 0, 0, EPI_BIAS, 0, 0},
      {"TN 872x2304x768 bias", 872, 2304, 768, 0, 0, 0, 0, EPI_BIAS, 0, 0},
      {"TN 872x3072x768 bias_gelu+C2", 872, 3072, 768, 0, 0, 0, 0, EPI_BIAS_GELU, 1, 0},
      {"TN 872x3072x768 gelu+gelu' (C2)", 872, 3072, 768, 0, 0, 0, 0, EPI_BIAS_GELU_GRAD, 1, 0},
      {"TN 300x192x136 gelu+gelu' tails", 300, 192, 136, 0, 0, 0, 0, EPI_BIAS_GELU_GRAD, 1, 0},
      {"NN dgrad 872x3072x768 mul", 872, 3072, 768, 0, 1, 0, 0, EPI_MUL, 0, 0},
      {"TN 872x768x3072 bias_resid", 872, 768, 3072, 0, 0, 0, 0, EPI_BIAS_RESID, 0, 0},
      {"TN 872x768x768 bias_resid drop", 872, 768, 768, 0, 0, 0, 0, EPI_BIAS_RESID, 0, 1},
      {"TN 64x768x768 bias_tanh", 64, 768, 768, 0, 0, 0, 0, EPI_BIAS_TANH, 0, 0},
      {"TN 40x30528x768 f32 out bias", 40, 30528, 768, 0, 0, 1, 0, EPI_BIAS, 0, 0},
      {"TN 8192x768x768 many tiles", 8192, 768, 768, 0, 0, 0, 0, EPI_NONE, 0, 0},
      {"NN dgrad 872x768x2304 resid", 872, 768, 2304, 0, 1, 0, 0, EPI_RESID, 0, 0},
      {"NN dgrad 872x3072x768 dgelu", 872, 3072, 768, 0, 1, 0, 0, EPI_DGELU, 0, 0},
      {"NN dgrad 300x192x136 none", 300, 192, 136, 0, 1, 0, 0, EPI_NONE, 0, 0},
      {"TT wgrad 768x768x872 acc", 768, 768, 872, 1, 1, 1, 1, EPI_NONE, 0, 0},
      {"TT wgrad 3072x768x8720 acc", 3072, 768, 8720, 1, 1, 1, 1, EPI_NONE, 0, 0},
      {"TT wgrad 768x2048x360 acc", 768, 2048, 360, 1, 1, 1, 1, EPI_NONE, 0, 0},
      {"TT 200x136x300 f32 store", 200, 136, 300, 1, 1, 1, 0, EPI_NONE, 0, 0},
  };
  for (int pair = 0; pair < 2 && !perf_only; ++pair) {
    const int bns[4] = {0, 128, 192, 256};
    for (int bi = 0; bi < 4; ++bi) {
      g_pair = pair; g_bn = bns[bi];
      printf("== tcgen05 GEMM, %s, N tile %s%d ==\n", pair ? "CTA pair (cta_group::2)" : "single CTA", g_bn ? "" : "auto ", g_bn);
      for (const Case& c : cases) {
        if (pair && c.b_mn && g_bn == 192) continue;      // MN-major B is loaded in 64-column groups per CTA
        fails += run_case(c, false, bi == 0 ? 40000 : 8000);
      }
    }
  }
  g_pair = -1; g_bn = 0;
  printf("== fp32 SIMT GEMM ==\n");
  const Case fcases[] = {
      {"TN 300x200x136 bias", 300, 200, 136, 0, 0, 0, 0, EPI_BIAS, 0, 0},
      {"TN 128x3072x768 bias_gelu+C2", 128, 3072, 768, 0, 0, 0, 0, EPI_BIAS_GELU, 1, 0},
      {"TN 128x768x768 bias_resid drop", 128, 768, 768, 0, 0, 0, 0, EPI_BIAS_RESID, 0, 1},
      {"NN 300x192x136 dgelu", 300, 192, 136, 0, 1, 0, 0, EPI_DGELU, 0, 0},
      {"TN 128x3072x768 gelu+gelu' (C2)", 128, 3072, 768, 0, 0, 0, 0, EPI_BIAS_GELU_GRAD, 1, 0},
      {"NN 300x192x136 mul", 300, 192, 136, 0, 1, 0, 0, EPI_MUL, 0, 0},
      {"TT 200x136x300 acc", 200, 136, 300, 1, 1, 1, 1, EPI_NONE, 0, 0},
  };
  for (const Case& c : fcases) if (!perf_only) fails += run_case(c, true, 40000);
  for (int pair = 0; pair < 2 && do_perf; ++pair) {
    g_pair = pair;
    printf("== perf (B=64, L=436: M=27904), %s ==\n", pair ? "CTA pair" : "single CTA");
    perf("QKV fwd", 27904, 2304, 768, 0, 0, 0, 0, EPI_BIAS);
    perf("out-proj fwd", 27904, 768, 768, 0, 0, 0, 0, EPI_BIAS_RESID);
    perf("FFN1 fwd", 27904, 3072, 768, 0, 0, 0, 0, EPI_BIAS_GELU);
    perf("FFN2 fwd", 27904, 768, 3072, 0, 0, 0, 0, EPI_BIAS_RESID);
    perf("FFN2 dgrad", 27904, 3072, 768, 0, 1, 0, 0, EPI_DGELU);
    perf("FFN1 dgrad", 27904, 768, 3072, 0, 1, 0, 0, EPI_RESID);
    perf("FFN1 wgrad", 3072, 768, 27904, 1, 1, 1, 1, EPI_NONE);
    perf("QKV wgrad", 2304, 768, 27904, 1, 1, 1, 1, EPI_NONE);
    perf("out wgrad", 768, 768, 27904, 1, 1, 1, 1, EPI_NONE);
  }
  printf("%s (%d failures)\n", fails ? "GEMM TEST FAILED" : "GEMM TEST PASSED", fails);
  return fails ? 1 : 0;
}
