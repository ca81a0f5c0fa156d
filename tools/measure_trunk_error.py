#!/usr/bin/env python
"""bf16 vs fp32 ResNet-50 trunk (cuDNN convolutions + the library's BatchNorm kernels, train-mode statistics) on 512x512
synthetic images: relative Frobenius error of the [B, 256, 2048] grid features.  Prints one JSON line."""
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
x = torch.randint(0, 256, (B, 3, 512, 512), dtype=torch.uint8, generator=torch.Generator().manual_seed(1)).cuda()
state = {k: v.clone() for k, v in enc.state_dict().items()}
f32 = enc.grid_features(x, dtype=torch.float32).double()
enc.load_state_dict(state)          # undo the running-statistics update
b16 = enc.grid_features(x, dtype=torch.bfloat16).double()
rel = float((b16 - f32).norm() / f32.norm())
mx = float((b16 - f32).abs().max() / f32.abs().max())
print(json.dumps({"what": "bf16 vs fp32 trunk, train-mode BN, random-init ResNet-50", "batch": B, "rel_l2": rel, "max_over_max": mx,
                  "feat_rms": float(f32.pow(2).mean().sqrt())}))
