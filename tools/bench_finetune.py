"""Report-generation fine-tune step throughput (BASELINE.json configs[4]) on one B200:
BERT-base, 256 regions + 253 report tokens (L = 512), s2s mask, max_pred = 10, bf16, dropout 0.1, BertAdam; synthetic
images / reports, random-init weights.  Prints one JSON line (samples/s).   python tools/bench_finetune.py [--batch 64]
"""
import argparse
import json
import os
import random
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import medvill_b200  # noqa: F401
from medvill_b200.config import BertConfig
from medvill_b200.report_generation import BertAdam, BertForPreTrainingLossMask, Preprocess4Seq2seq

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:       # one process per GPU (torchrun): gradients averaged over ranks as the reference's DistributedDataParallel does
    torch.distributed.init_process_group("nccl", device_id=dev)
B = a.batch
args = types.SimpleNamespace(img_hidden_sz=2048, hidden_size=768, img_postion=True, img_encoding="fully_use_cnn", allow_random_trunk=True, len_vis_input=256,
                             img_size=512, max_len_b=253, precision="bf16", max_micro_batch=B, tasks="report_generation")
torch.manual_seed(0)
model = BertForPreTrainingLossMask(BertConfig.from_pretrained("bert-base-uncased"), args, len_vis_input=256).to(dev).train()
if world > 1:
    model.init_distributed()
words = ["[PAD]"] + ["w%d" % i for i in range(1, 30522)]
for tok, i in (("[UNK]", 100), ("[CLS]", 101), ("[SEP]", 102), ("[MASK]", 103)):
    words[i] = tok
stoi = {w: i for i, w in enumerate(words)}
pipe = Preprocess4Seq2seq(args, 10, 0.15, words, lambda t: [stoi[x] for x in t], 512, False, mode="s2s", len_vis_input=256,
                          truncate_config={"max_len_b": 253, "trunc_seg": "b", "always_truncate_tail": False}, compact_mask=True,
                          image_loader=lambda p: None)
random.seed(123 + rank)
rng = np.random.RandomState(123 + rank)
rows = [pipe((None, [words[t] for t in rng.randint(999, 30522, size=rng.randint(16, 254))], None, None, None)) for _ in range(B)]
col = lambda i: torch.tensor([r[i] for r in rows])
input_ids, segment_ids, masked_ids, masked_pos, masked_w = col(0).to(dev), col(1).to(dev), col(3), col(4), col(5)
cm = torch.stack([r[2] for r in rows])
mode, t_len = cm[:, 0].to(torch.uint8).to(dev), cm[:, 1].to(torch.int32).to(dev)
img = torch.randint(0, 256, (B, 3, 512, 512), dtype=torch.uint8, device=dev)
opt = BertAdam([{"params": list(model.parameters()), "weight_decay": 0.01}], lr=3e-5, warmup=0.1, t_total=1000, model=model)


def run(n):
    pending, out = None, None
    for _ in range(n):
        nxt = model.finetune_step(img, input_ids, segment_ids, None, masked_ids, masked_pos, masked_w, optimizer=opt, mode=mode,
                                  t_len=t_len, lazy=True)
        if pending is not None:
            out = pending()
        pending = nxt
    return pending()


run(a.warmup)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = run(a.steps)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t)
if rank == 0:
  print(json.dumps({"metric": "report-generation fine-tune samples/sec", "value": B * world / (ms / 1e3), "unit": "samples/s", "n_gpus": world,
                  "ms_per_step": ms, "batch_per_gpu": B, "joint_len": 512, "max_pred": 10, "mask": "s2s (fine-tune variant)",
                  "optimizer": "BertAdam (mv_bert_adam_step)", "encoder_tflops_dense_per_gpu": B * 2.8991e11 / (ms / 1e3) / 1e12,
                  "last_loss": out["loss"], "dtype": "bf16", "data": "synthetic"}))
if world > 1:
    torch.cuda.synchronize()
    torch.distributed.barrier()
    model._cxrbert._release_engine()
    torch.distributed.destroy_process_group()
