// gemm.h — host-side contract shared by the tcgen05 GEMM (bf16 operands, production) and the SIMT fp32
// GEMM (check mode). One descriptor covers the three contractions of a linear layer:
//   forward  Y[M,N]  = X[M,K]  · W[N,K]^T          a_mn=0, b_mn=0   (reference: nn.Linear, upstream BertSelfAttention
//                                                                    twin Downstream_task/report_generation_and_vqa/sc/
//                                                                    pytorch_pretrained_bert/model.py:273-298)
//   dgrad    dX[M,K] = dY[M,N] · W[N,K]            a_mn=0, b_mn=1   (W read in place: its K_in axis is contiguous)
//   wgrad    dW[N,K] = dY[M,N]^T · X[M,K]          a_mn=1, b_mn=1   (both activations read in place, fp32 reduce-add)
#pragma once
#include "common.cuh"

namespace mv {

enum Epilogue : int {
  EPI_NONE = 0,        // C = acc
  EPI_BIAS = 1,        // C = acc + bias[n]
  EPI_BIAS_GELU = 2,   // C2 = acc + bias (pre-activation, optional); C = gelu_erf(C2)
  EPI_BIAS_RESID = 3,  // C = dropout(acc + bias) + resid[m,n]
  EPI_BIAS_TANH = 4,   // C = tanh(acc + bias)
  EPI_RESID = 5,       // C = acc + resid[m,n]
  EPI_DGELU = 6,       // C = acc * gelu_erf'(aux[m,n])
  EPI_BIAS_GELU_GRAD = 7,  // C = gelu_erf(acc + bias); C2 = gelu_erf'(acc + bias): the derivative is saved INSTEAD of the pre-activation,
                           // so the backward GEMM's epilogue is a plain multiply (EPI_MUL) rather than a second erf evaluation
  EPI_MUL = 8,         // C = acc * aux[m,n]
};

struct GemmDesc {
  int M = 0, N = 0, K = 0;
  const void* A = nullptr; long lda = 0; int a_mn = 0;  // a_mn=0: [M,K] rows; a_mn=1: stored [K,M] rows
  const void* B = nullptr; long ldb = 0; int b_mn = 0;  // b_mn=0: [N,K] rows; b_mn=1: stored [K,N] rows
  void* C = nullptr; long ldc = 0; int c_f32 = 0;       // output dtype: activation dtype, or fp32
  int accumulate = 0;                                   // fp32 outputs only: C += result (split-K reduce-add)
  void* C2 = nullptr; long ldc2 = 0;                    // optional pre-activation output (activation dtype)
  int epi = EPI_NONE;
  const float* bias = nullptr;                          // [N] fp32 (master parameter, never shadowed)
  const void* resid = nullptr; long ldr = 0;            // [M,N] activation dtype, or fp32 when resid_f32 (then c_f32 too)
  int resid_f32 = 0;                                    // fp32 residual stream: resid AND C are fp32 (EPI_BIAS_RESID only)
  const void* aux = nullptr; long ldaux = 0;            // [M,N] activation dtype
  int drop_on = 0; uint32_t drop_site = 0; DropoutCfg drop = {0.f, 0u, 0u, 1.f, 0ull};
  int splitk = 0;                                       // 0 = auto (only used when accumulate=1)
  int pair = -1;                                        // CTA-pair (cta_group::2, 256-row tiles): -1 auto, 0 off, 1 on
  int bn = 0;                                           // N tile: 0 auto, else 128 / 192 / 256 (tests, tuning)
  // Sliding-window A operand (tcgen05 path, a_mn = 0): row m = (b, y, x) of an output image [win_b, win_ho, win_wo]; k-block kb
  // (64 elements) is the contiguous run of 64 input elements starting at pixel (y + kb, x) of the channels-last input
  // A[win_b, win_hs, win_ws, win_c] (win_c channels per pixel: 64 / win_c pixels per run).  That is a convolution with a
  // (K / 64) x (64 / win_c) window, stride 1, no padding, read in place through a 4-D tensor map whose pixel stride
  // (win_c elements) is smaller than the 64-element box: no im2col copy.  M = win_b * win_ho * win_wo, win_wo % 128 == 0.
  int a_win = 0, win_b = 0, win_ho = 0, win_wo = 0, win_hs = 0, win_ws = 0, win_c = 0;
};

// bf16 operands / fp32 accumulate in TMEM / bf16 or fp32 output. sm_100a only.
int gemm_bf16_tc05(const GemmDesc& d, cudaStream_t stream);
// fp32 operands, CUDA-core FMA (check mode, 1e-4 parity gate).
int gemm_f32_simt(const GemmDesc& d, cudaStream_t stream);

}  // namespace mv
