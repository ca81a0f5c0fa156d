#!/usr/bin/env python
"""bf16 vs fp32 ResNet-50 trunk (cuDNN convolutions + the library's BatchNorm kernels, train-mode statistics) on 512x512
synthetic images: relative Frobenius error of the [B, 256, 2048] grid features, next to the same comparison for PyTorch's own
bf16 autocast of the torchvision module (the floor any bf16 trunk has on these weights), for
  * random-init weights (torchvision default init: a BatchNorm'd random ReLU network amplifies perturbations layer by layer), and
  * the same weights with every block's last BatchNorm scale damped (bn3.weight = 0.2), a well-conditioned network in the way a
    trained one is (zero_init_residual-style).
Prints one JSON line per weight set."""
import json
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MEDVILL_ALLOW_RANDOM_TRUNK", "1")
import medvill_b200  # noqa: E402,F401
from medvill_b200.models.image import ImageEncoder_cnn  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
enc = ImageEncoder_cnn(types.SimpleNamespace(num_image_embeds=180, allow_random_trunk=True)).cuda().train()
x8 = torch.randint(0, 256, (B, 3, 512, 512), dtype=torch.uint8, generator=torch.Generator().manual_seed(1)).cuda()
mean = torch.tensor([0.485, 0.456, 0.406], device="cuda").view(1, 3, 1, 1)
std = torch.tensor([0.229, 0.224, 0.225], device="cuda").view(1, 3, 1, 1)
xf = (x8.float() / 255.0 - mean) / std


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


for label, gamma3 in (("random-init", None), ("bn3.weight=0.2", 0.2)):
    if gamma3 is not None:
        with torch.no_grad():
            for n, p in enc.named_parameters():
                if n.endswith("bn3.weight"):
                    p.fill_(gamma3)
        enc._exec = None
    state = {k: v.clone() for k, v in enc.state_dict().items()}

    def run(fn):
        enc.load_state_dict(state)          # same running statistics before every run
        with torch.no_grad():
            return fn()

    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        ref = run(lambda: enc.model(xf)).permute(0, 2, 3, 1).reshape(B, -1, 2048)          # torchvision fp32, train-mode BN
    f32 = run(lambda: enc.grid_features(x8, dtype=torch.float32))
    b16 = run(lambda: enc.grid_features(x8, dtype=torch.bfloat16))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ac = run(lambda: enc.model(xf)).permute(0, 2, 3, 1).reshape(B, -1, 2048)
    enc.load_state_dict(state)
    print(json.dumps({"weights": label, "batch": B, "ours_fp32_vs_torch_fp32": rel(f32, ref), "ours_bf16_vs_torch_fp32": rel(b16, ref),
                      "torch_autocast_bf16_vs_torch_fp32": rel(ac, ref), "feat_rms": float(ref.pow(2).mean().sqrt())}))
