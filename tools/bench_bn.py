"""BatchNorm kernels of the trunk at the shapes of one B = 64 pre-training step: achieved HBM GB/s of the statistics and the
apply pass (CUDA events, 20 launches each).  python tools/bench_bn.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import medvill_b200  # noqa: F401
from medvill_b200 import _lib

L = _lib
dev = "cuda:0"
shapes = [(64 * 128 * 128, 64, 0), (64 * 128 * 128, 256, 1), (64 * 64 * 64, 128, 0), (64 * 64 * 64, 512, 1), (64 * 32 * 32, 256, 0),
          (64 * 32 * 32, 1024, 1), (64 * 16 * 16, 512, 0), (64 * 16 * 16, 2048, 1)]
tot = {"train": [0.0, 0.0], "eval": [0.0, 0.0]}
for rows, C, res in shapes:
    x = torch.randn(rows, C, device=dev).to(torch.bfloat16)
    r = torch.randn(rows, C, device=dev).to(torch.bfloat16) if res else None
    y = torch.empty_like(x)
    w, b, rm, rv = torch.ones(C, device=dev), torch.zeros(C, device=dev), torch.zeros(C, device=dev), torch.ones(C, device=dev)
    ws = torch.zeros(int(L.lib().mv_bn_workspace_floats(rows, C)), dtype=torch.float32, device=dev)

    def run(training):
        L.check(L.lib().mv_bn_forward(L.ptr(x), L.ptr(r), L.ptr(y), rows, C, L.ptr(w), L.ptr(b), L.ptr(rm), L.ptr(rv), 0.1, 1e-5, training, 1,
                                      L.ptr(ws), ws.numel(), L.MV_PREC_BF16, L.stream_ptr()))
    out = {}
    for name, training in (("train", 1), ("eval", 0)):
        for _ in range(3):
            run(training)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run(training)
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 20
    nbytes_apply = rows * C * 2 * (3 if res else 2)
    nbytes_stats = rows * C * 2
    stats_ms = out["train"] - out["eval"]
    print("rows %8d C %4d resid %d: apply %.3f ms (%.0f GB/s)  stats+finalize %.3f ms (%.0f GB/s)" % (
        rows, C, res, out["eval"], nbytes_apply / out["eval"] / 1e6, stats_ms, nbytes_stats / stats_ms / 1e6))
