"""Visual encoder: ResNet-50 trunk -> grid-region features -> random region sampling.

Mirrors /root/reference/models/image.py:46-69 (`ImageEncoder_cnn`; exported also as `ImageEncoder`, the name the
north_star and Classification/mmbt/models/image.py:16 use).  The convolutions stay on cuDNN through torchvision
(BASELINE.json north_star (4)): frozen weights, train-mode BatchNorm (models/train_origin.py:72), channels-last so that
the [B, 2048, g, g] feature map *is* the [B, g*g, 2048] region matrix with no flatten/transpose copies; region gather +
2048->768 projection + position/type add + LayerNorm run in libmedvill_sm100 (mv_forward).
"""
import torch
import torch.nn as nn
import torchvision


class ImageEncoder_cnn(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        try:
            model = torchvision.models.resnet50(weights="IMAGENET1K_V1")      # reference: pretrained=True (image.py:50)
        except Exception:                                                     # offline: random init, same architecture
            model = torchvision.models.resnet50(weights=None)
        self.model = nn.Sequential(*list(model.children())[:-2])
        self.region_idx_override = None     # parity runs inject the sampled regions (the reference uses the CPU RNG)

    def grid_features(self, x, dtype=None):
        """[B, 3, h, w] -> [B, (h/32)*(w/32), 2048], contiguous, channels-last under the hood."""
        if dtype is not None and x.dtype != dtype:
            x = x.to(dtype)
        if x.is_cuda:
            x = x.contiguous(memory_format=torch.channels_last)
        out = self.model(x)                                    # [B, 2048, g, g]
        B, Cc = out.shape[0], out.shape[1]
        return out.permute(0, 2, 3, 1).reshape(B, -1, Cc)      # free view when `out` is channels-last

    def sample_regions(self, num_range):
        """models/image.py:63-65 — one permutation per forward, shared by the batch, sorted."""
        if self.region_idx_override is not None:
            return torch.as_tensor(self.region_idx_override, dtype=torch.long)
        idx = torch.randperm(num_range)[:self.args.num_image_embeds]
        idx, _ = torch.sort(idx)
        return idx

    def forward(self, x):
        out = self.grid_features(x)
        idx = self.sample_regions(out.size(1)).to(out.device)
        vis_pe = torch.arange(out.size(1), dtype=torch.long, device=out.device).unsqueeze(0).expand(out.size(0), out.size(1))
        return out[:, idx], vis_pe[:, idx]


ImageEncoder = ImageEncoder_cnn
