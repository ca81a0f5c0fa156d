"""Stem convolution at the bench shape (B = 64, 512 x 512 images -> s2d [64, 259, 259, 16]): mv_stem_conv_s2d (tcgen05 GEMM with a
sliding-window A operand) next to cuDNN's implicit GEMM on the same tensors.  CUDA events, 20 launches.  python tools/bench_stem.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import medvill_b200  # noqa: F401
from medvill_b200 import _lib

B, Hs = 64, 259
x = torch.randn(B, 16, Hs, Hs, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
w = (torch.randn(64, 16, 4, 4, device="cuda") * 0.1).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
y = torch.empty(B, 64, Hs - 3, Hs - 3, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)


def ours():
    _lib.check(_lib.lib().mv_stem_conv_s2d(_lib.ptr(x), _lib.ptr(w), _lib.ptr(y), B, Hs, Hs, 64, _lib.stream_ptr(x.device)), "stem")


def cudnn():
    return F.conv2d(x, w)


flop = 2.0 * B * 256 * 256 * 64 * 256
nbytes = x.numel() * 2 + y.numel() * 2
for name, fn in (("mv_stem_conv_s2d", ours), ("cuDNN conv2d", cudnn)):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("%-18s %.3f ms  %.0f TFLOP/s  %.0f GB/s (input + output once)" % (name, ms, flop / ms / 1e9, nbytes / ms / 1e6))
