// gemm_simt.cu — fp32 "check mode" GEMM on CUDA cores (FFMA, fp32 operands and accumulation). It exists so the
// whole pre-training step can be verified against the CPU oracle at 1e-4 (BASELINE.json north_star: "1e-4 for
// an fp32 check mode"); it shares GemmDesc and the exact epilogue semantics with the tcgen05 kernel.
#include "gemm.h"

namespace mv {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct SimtParams {
  int M, N, K;
  const float* A; long sam, sak;
  const float* B; long sbn, sbk;
  float* C; long ldc;
  float* C2; long ldc2;
  int epi, accumulate;
  const float* bias;
  const float* resid; long ldr;
  const float* aux; long ldaux;
  int drop_on; uint32_t drop_site; DropoutCfg drop;
};

__global__ void __launch_bounds__(256) gemm_f32_kernel(const SimtParams p) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < p.K; k0 += TK) {
    for (int i = threadIdx.x; i < TM * TK; i += 256) {
      int mm, kk;
      if (p.sak == 1) { kk = i % TK; mm = i / TK; } else { mm = i % TM; kk = i / TM; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < p.M && k < p.K) ? p.A[m * p.sam + k * p.sak] : 0.f;
    }
    for (int i = threadIdx.x; i < TN * TK; i += 256) {
      int nn, kk;
      if (p.sbk == 1) { kk = i % TK; nn = i / TK; } else { nn = i % TN; kk = i / TN; }
      const int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < p.N && k < p.K) ? p.B[n * p.sbn + k * p.sbk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      const int epi = p.epi;
      if (epi == EPI_BIAS || epi == EPI_BIAS_GELU || epi == EPI_BIAS_RESID || epi == EPI_BIAS_TANH || epi == EPI_BIAS_GELU_GRAD) v += p.bias[n];
      if (epi == EPI_BIAS_GELU) {
        if (p.C2) p.C2[m * p.ldc2 + n] = v;
        v = gelu_erf(v);
      } else if (epi == EPI_BIAS_GELU_GRAD) {
        if (p.C2) p.C2[m * p.ldc2 + n] = gelu_erf_grad(v);
        v = gelu_erf(v);
      } else if (epi == EPI_BIAS_TANH) {
        v = tanhf(v);
      } else if (epi == EPI_BIAS_RESID || epi == EPI_RESID) {
        if (p.drop_on && epi == EPI_BIAS_RESID) {
          const uint64_t e = static_cast<uint64_t>(m) * p.N + n;
          const uint4 keep = dropout_keep16(p.drop, p.drop_site, e >> 4);
          v = keep16_bit(keep, static_cast<int>(e & 15)) ? v * p.drop.scale : 0.f;
        }
        v += p.resid[m * p.ldr + n];
      } else if (epi == EPI_DGELU) {
        v *= gelu_erf_grad(p.aux[m * p.ldaux + n]);
      } else if (epi == EPI_MUL) {
        v *= p.aux[m * p.ldaux + n];
      }
      float* dst = p.C + m * p.ldc + n;
      if (p.accumulate) *dst += v; else *dst = v;
    }
  }
}

}  // namespace

int gemm_f32_simt(const GemmDesc& d, cudaStream_t stream) {
  MV_REQUIRE(d.M > 0 && d.N > 0 && d.K > 0, "gemm_f32: empty problem M=%d N=%d K=%d", d.M, d.N, d.K);
  MV_REQUIRE(d.A && d.B && d.C, "gemm_f32: null operand");
  SimtParams p;
  p.M = d.M; p.N = d.N; p.K = d.K;
  p.A = static_cast<const float*>(d.A);
  if (!d.a_mn) { p.sam = d.lda; p.sak = 1; } else { p.sam = 1; p.sak = d.lda; }
  p.B = static_cast<const float*>(d.B);
  if (!d.b_mn) { p.sbn = d.ldb; p.sbk = 1; } else { p.sbn = 1; p.sbk = d.ldb; }
  p.C = static_cast<float*>(d.C); p.ldc = d.ldc;
  p.C2 = static_cast<float*>(d.C2); p.ldc2 = d.ldc2;
  p.epi = d.epi; p.accumulate = d.accumulate;
  p.bias = d.bias;
  p.resid = static_cast<const float*>(d.resid); p.ldr = d.ldr;
  p.aux = static_cast<const float*>(d.aux); p.ldaux = d.ldaux;
  p.drop_on = d.drop_on; p.drop_site = d.drop_site; p.drop = d.drop;
  dim3 grid((d.N + TN - 1) / TN, (d.M + TM - 1) / TM);
  gemm_f32_kernel<<<grid, 256, 0, stream>>>(p);
  MV_LAUNCH_CHECK();
  return 0;
}

}  // namespace mv
