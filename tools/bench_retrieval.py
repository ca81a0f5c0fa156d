"""Label-conditioned ITM retrieval scoring throughput (BASELINE.json configs[4]) on one B200:
BERT-base, N=180 regions, S=253 (L=436), bf16, synthetic images / reports / 14-class label sets (aligned iff the label
sets are equal).  ResNet-50 features are computed once per image; the [images x reports] similarity matrix is filled on
the device by mv_forward + mv_itm_match_prob.  Prints one JSON line (pairs/s, image features/s, rank metrics).

  python tools/bench_retrieval.py [--images 256] [--reports 256] [--pair-batch 64]
The full 3k x 3k problem is 9.0 M pair forwards (~81 GFLOP each): --images 3000 --reports 3000 (about 15 min on one GPU).
"""
import argparse
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import medvill_b200  # noqa: F401
from medvill_b200.config import BertConfig
from medvill_b200.retrieval import CXRBertForRetrieval, RetrievalScorer, evaluate

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=256)
ap.add_argument("--reports", type=int, default=256)
ap.add_argument("--pair-batch", type=int, default=64)
a = ap.parse_args()
dev = torch.device("cuda:0")
margs = types.SimpleNamespace(img_hidden_sz=2048, embedding_size=768, hidden_size=768, dropout_prob=0.1, img_encoder="random-pixel", allow_random_trunk=True,
                              num_image_embeds=180, img_size=512, seq_len=253, lr=1e-5, precision="bf16", max_micro_batch=a.pair_batch,
                              seed=123, weight_load=False)
torch.manual_seed(0)
model = CXRBertForRetrieval(BertConfig.from_pretrained("bert-base-uncased"), margs).to(dev).eval()
rng = np.random.RandomState(123)
images = torch.randint(0, 256, (a.images, 3, 512, 512), dtype=torch.uint8)
t_len = rng.randint(17, 255, size=a.reports).astype(np.int32)          # real text length incl. [SEP]
ids = np.zeros((a.reports, 254), dtype=np.int64)
for j, t in enumerate(t_len):
    ids[j, :t - 1] = rng.randint(999, 30522, size=t - 1)
    ids[j, t - 1] = 102
img_lab = rng.randint(0, 14, size=a.images)
txt_lab = rng.randint(0, 14, size=a.reports)
scorer = RetrievalScorer(model, pair_batch=a.pair_batch)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
scorer.image_features(images[:8])                                        # warm-up (cuDNN plans, workspaces)
scorer.score_matrix(scorer.image_features(images[:2]), torch.from_numpy(ids[:64]), torch.from_numpy(t_len[:64]))
torch.cuda.synchronize()
ev[0].record()
feats = scorer.image_features(images)
ev[1].record()
sims = scorer.score_matrix(feats, torch.from_numpy(ids), torch.from_numpy(t_len))
ev[2].record()
host = sims.cpu().numpy()
torch.cuda.synchronize()
ms_feat, ms_pairs = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
labels = (img_lab[:, None] == txt_lab[None, :]).astype(np.int64)
args = types.SimpleNamespace(eval_len_size=a.reports, i2t=True, t2i=False)
res, aligned, mrr, rp = evaluate(args, list(host.reshape(-1)), labels.reshape(-1).tolist(), list(range(a.images * a.reports)))
pairs = a.images * a.reports
print(json.dumps({"metric": "ITM retrieval scoring", "value": pairs / (ms_pairs / 1e3), "unit": "pairs/s", "n_gpus": 1,
                  "images": a.images, "reports": a.reports, "pair_batch": a.pair_batch, "ms_pairs": ms_pairs,
                  "image_features_per_s": a.images / (ms_feat / 1e3), "ms_image_features": ms_feat,
                  "encoder_tflops_dense": pairs * 8.1071e10 / (ms_pairs / 1e3) / 1e12,
                  "projected_3k_x_3k_minutes": 9e6 / (pairs / (ms_pairs / 1e3)) / 60.0,
                  "hit@1/5/10": [res["i2t_retrieval"][k] for k in ("R@1", "R@5", "R@10")], "mrr": float(mrr),
                  "dtype": "bf16", "data": "synthetic (random-init weights: metrics are chance level by construction)"}))
