"""Calibration only (never on the product path): cuBLAS bf16 throughput on the GEMM shapes of one encoder layer at B=64,
L=436, to know what the tensor pipe + memory system of this box can deliver for them.  Random operands."""
import torch

torch.manual_seed(0)
dev = "cuda"
M = 27904
shapes = [("QKV fwd", M, 2304, 768, "tn"), ("out-proj fwd", M, 768, 768, "tn"), ("FFN1 fwd", M, 3072, 768, "tn"), ("FFN2 fwd", M, 768, 3072, "tn"),
          ("FFN2 dgrad", M, 3072, 768, "nn"), ("FFN1 dgrad", M, 768, 3072, "nn"), ("FFN1 wgrad", 3072, 768, M, "tt"),
          ("QKV wgrad", 2304, 768, M, "tt"), ("out wgrad", 768, 768, M, "tt")]
for name, m, n, k, kind in shapes:
    if kind == "tn":
        a = torch.randn(m, k, device=dev, dtype=torch.bfloat16); b = torch.randn(n, k, device=dev, dtype=torch.bfloat16)
        f = lambda: a @ b.t()
    elif kind == "nn":
        a = torch.randn(m, k, device=dev, dtype=torch.bfloat16); b = torch.randn(k, n, device=dev, dtype=torch.bfloat16)
        f = lambda: a @ b
    else:
        a = torch.randn(k, m, device=dev, dtype=torch.bfloat16); b = torch.randn(k, n, device=dev, dtype=torch.bfloat16)
        f = lambda: a.t() @ b
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("  cublas %-14s M=%d N=%d K=%d: %.3f ms  %.1f TFLOP/s" % (name, m, n, k, ms, 2.0 * m * n * k / ms / 1e9))
