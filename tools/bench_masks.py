"""Attention-mask sweep (BASELINE.json configs[2]): fused masked attention forward + backward on one B200 at joint length 512
(256 regions + 253 report tokens), B=64, 12 heads, bf16, dropout 0.1, for Bidirectional, Seq2Seq, Bi&Seq2Seq mixed (75/25 per
sample), Non-cross and Bidirectional Auto-Regressive.  Prints one JSON line per mode: ms per layer, dense TFLOP/s (4 L^2 d
forward, 10 L^2 d backward per head) and the mask density (fraction of the L x L score matrix that is not masked; tiles that
are entirely masked are skipped by the kernels).   python tools/bench_masks.py [--L 512] [--batch 64]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import medvill_b200  # noqa: F401
from medvill_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=512)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
B, nh, Lq = a.batch, 12, a.L
A = Lq - 254
H = nh * 64
dev = "cuda:0"
rng = np.random.RandomState(123)
t_len = torch.tensor(rng.randint(17, 255, size=B), dtype=torch.int32, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(B, Lq, 3 * H, device=dev, generator=g).to(torch.bfloat16)
dctx = torch.randn(B, Lq, H, device=dev, generator=g).to(torch.bfloat16)
ctx = torch.empty(B, Lq, H, device=dev, dtype=torch.bfloat16)
dqkv = torch.empty(B, Lq, 3 * H, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, nh, Lq, device=dev)
delta = torch.empty(B, nh, Lq, device=dev)
dq_acc = torch.empty(B * Lq, H, device=dev)
modes = {"bidirectional": [L.MODE_BIDIR] * B, "seq2seq": [L.MODE_S2S] * B,
         "mixed_75_25": [L.MODE_S2S if rng.rand() < 0.75 else L.MODE_BIDIR for _ in range(B)],
         "non_cross": [L.MODE_NONCROSS] * B, "bidirectional_auto_regressive": [L.MODE_BAR] * B}
for name, ml in modes.items():
    mode = torch.tensor(ml, dtype=torch.uint8, device=dev)
    mask = torch.empty(B, Lq, Lq, dtype=torch.uint8, device=dev)
    L.check(L.lib().mv_attn_mask_dump(L.ptr(mode), L.ptr(t_len), B, A, Lq, L.ptr(mask), L.stream_ptr()))
    density = float(mask.float().mean())

    def fwd():
        L.check(L.lib().mv_attention_fwd(B, Lq, nh, A, L.ptr(mode), L.ptr(t_len), L.ptr(qkv), L.ptr(ctx), L.ptr(lse), 0.1, 7, 16,
                                         L.MV_PREC_BF16, L.stream_ptr()))

    def bwd():
        L.check(L.lib().mv_attention_bwd(B, Lq, nh, A, L.ptr(mode), L.ptr(t_len), L.ptr(qkv), L.ptr(ctx), L.ptr(lse), L.ptr(dctx),
                                         L.ptr(dqkv), L.ptr(dq_acc), L.ptr(delta), 0.1, 7, 16, L.MV_PREC_BF16, L.stream_ptr()))

    for _ in range(3):
        fwd(); bwd()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(a.iters):
        fwd()
    ev[1].record()
    for _ in range(a.iters):
        bwd()
    ev[2].record()
    torch.cuda.synchronize()
    f_ms, b_ms = ev[0].elapsed_time(ev[1]) / a.iters, ev[1].elapsed_time(ev[2]) / a.iters
    fl = 4.0 * B * nh * Lq * Lq * 64
    print(json.dumps({"mask": name, "joint_len": Lq, "batch": B, "density": round(density, 4), "fwd_ms": f_ms, "bwd_ms": b_ms,
                      "fwd_tflops_dense": fl / f_ms / 1e9, "bwd_tflops_dense": 2.5 * fl / b_ms / 1e9,
                      "fwd_tflops_mask_aware": fl * density / f_ms / 1e9, "bwd_tflops_mask_aware": 2.5 * fl * density / b_ms / 1e9,
                      "dtype": "bf16", "dropout": 0.1}))
