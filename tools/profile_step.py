"""One warm-up step, then exactly one profiled pre-training step (cudaProfilerStart/Stop) — the target of
`ncu --profile-from-start off ...` launch lists and `--set full` captures (profiles/)."""
import argparse
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import medvill_b200  # noqa: F401
from medvill_b200.config import BertConfig
from medvill_b200.data.synthetic import synthetic_batch
from medvill_b200.models import CXRBERT

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--dropout", type=float, default=0.1)
ap.add_argument("--warmup", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda:0")
margs = types.SimpleNamespace(img_hidden_sz=2048, embedding_size=768, hidden_size=768, dropout_prob=a.dropout, img_encoder="random-pixel", allow_random_trunk=True,
                              num_image_embeds=180, img_size=512, seq_len=253, lr=1e-5, precision="bf16", max_micro_batch=a.batch, seed=123)
torch.manual_seed(0)
model = CXRBERT(BertConfig.from_pretrained("bert-base-uncased"), margs).to(dev).train()
b = synthetic_batch(a.batch, seed=123, image_dtype=torch.uint8)      # uint8 pixels on the wire, as bench.py
d = {k: (v if k == "txt_labels" else v.to(dev)) for k, v in b.items()}
step = lambda: model.pretrain_step(d["cls_tok"], d["input_ids"], d["txt_labels"], None, d["image"], d["segment"], d["is_aligned"],
                                   d["sep_tok"], mode=d["mode"], t_len=d["t_len"])
for _ in range(a.warmup):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss %.5f" % out["loss"])
