"""Label-conditioned ITM retrieval scoring (BASELINE.json configs[4]) on N B200s:
BERT-base, N=180 regions, S=253 (L=436), bf16, synthetic images / reports / 14-class label sets (aligned iff the label
sets are equal).  ResNet-50 features are computed once per image; the [images x reports] similarity matrix is filled on
the device by mv_forward + mv_itm_match_prob.  The image ROWS are sharded over the ranks (no data-path collective: the
scores of a pair depend on that pair only); rank 0 gathers the row blocks for the rank metrics.  Prints one JSON line
(pairs/s over all ranks from the slowest rank's device time, rank metrics, and the deviation of a small sub-block of the
similarity matrix from the CPU oracle on the same weights / grid features).

  python tools/bench_retrieval.py [--images 256] [--reports 256] [--pair-batch 64] [--oracle-block 4]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_retrieval.py --images 3000 --reports 3000
"""
import argparse
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import medvill_b200  # noqa: F401
from medvill_b200.config import BertConfig
from medvill_b200.retrieval import CXRBertForRetrieval, RetrievalScorer, evaluate

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=256)
ap.add_argument("--reports", type=int, default=256)
ap.add_argument("--pair-batch", type=int, default=64)
ap.add_argument("--oracle-block", type=int, default=4, help="side of the sub-block re-scored by the CPU oracle on rank 0 (0 = skip)")
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=dev)
margs = types.SimpleNamespace(img_hidden_sz=2048, embedding_size=768, hidden_size=768, dropout_prob=0.1, img_encoder="random-pixel", allow_random_trunk=True,
                              num_image_embeds=180, img_size=512, seq_len=253, lr=1e-5, precision="bf16", max_micro_batch=a.pair_batch,
                              seed=123, weight_load=False)
torch.manual_seed(0)                                                     # identical weights on every rank
model = CXRBertForRetrieval(BertConfig.from_pretrained("bert-base-uncased"), margs).to(dev).eval()
rng = np.random.RandomState(123)
t_len = rng.randint(17, 255, size=a.reports).astype(np.int32)          # real text length incl. [SEP]
ids = np.zeros((a.reports, 254), dtype=np.int64)
for j, t in enumerate(t_len):
    ids[j, :t - 1] = rng.randint(999, 30522, size=t - 1)
    ids[j, t - 1] = 102
img_lab = rng.randint(0, 14, size=a.images)
txt_lab = rng.randint(0, 14, size=a.reports)
# this rank's block of image rows; image i is generated from its own seed, so any sharding sees the same data
r0, r1 = a.images * rank // world, a.images * (rank + 1) // world


def image_block(lo, hi):
    out = torch.empty(hi - lo, 3, 512, 512, dtype=torch.uint8)
    for i in range(lo, hi):
        out[i - lo] = torch.randint(0, 256, (3, 512, 512), dtype=torch.uint8, generator=torch.Generator().manual_seed(10_000 + i))
    return out


images = image_block(r0, r1)
scorer = RetrievalScorer(model, pair_batch=a.pair_batch)
regions = torch.sort(torch.randperm(256, generator=torch.Generator().manual_seed(7))[:180]).values      # one region draw for the run
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
scorer.image_features(images[:min(8, len(images))])                                        # warm-up (cuDNN plans, workspaces)
scorer.score_matrix(scorer.image_features(images[:2]), torch.from_numpy(ids[:64]), torch.from_numpy(t_len[:64]), region_idx=regions)
torch.cuda.synchronize()
if world > 1:
    torch.distributed.barrier()
t_wall = time.perf_counter()
ev[0].record()
feats = scorer.image_features(images)
ev[1].record()
sims = scorer.score_matrix(feats, torch.from_numpy(ids), torch.from_numpy(t_len), region_idx=regions)
ev[2].record()
torch.cuda.synchronize()
ms = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])], dtype=torch.float64, device=dev)
if world > 1:
    torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)                    # the slowest rank sets the time
    blocks = [torch.empty(a.images * (r + 1) // world - a.images * r // world, a.reports, dtype=torch.float32, device=dev) for r in range(world)]
    torch.distributed.all_gather(blocks, sims.contiguous())
    full = torch.cat(blocks, 0)
else:
    full = sims
wall = time.perf_counter() - t_wall
ms_feat, ms_pairs = float(ms[0]), float(ms[1])
if rank == 0:
    host = full.cpu().numpy()
    labels = (img_lab[:, None] == txt_lab[None, :]).astype(np.int64)
    args = types.SimpleNamespace(eval_len_size=a.reports, i2t=True, t2i=False)
    res, aligned, mrr, rp = evaluate(args, list(host.reshape(-1)), labels.reshape(-1).tolist(), list(range(a.images * a.reports)))
    oracle = None
    nb = min(a.oracle_block, r1 - r0, a.reports)
    if nb > 0:
        # the same sub-block on the CPU oracle: same weights (state_dict -> reference names), same grid features, same regions
        from oracle import medvill_oracle as orc

        cfg = orc.Cfg()
        sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
        sd["mlm.predictions.bias"] = torch.zeros(cfg.vocab)             # (unused by the retrieval path)
        f = feats[:nb].float().cpu()
        want = np.zeros((nb, nb), dtype=np.float64)
        for i in range(nb):
            L = cfg.L
            masks = np.zeros((nb, L), dtype=np.int64)
            for j in range(nb):
                masks[j, :cfg.A + int(t_len[j])] = 1
            b = dict(cls_tok=np.full((nb, 1), 101, dtype=np.int64), sep_tok=np.full((nb, 1), 102, dtype=np.int64), input_ids=ids[:nb],
                     segment=np.ones((nb, 254), dtype=np.int64), attn_masks=masks, region_idx=regions.numpy())
            with torch.no_grad():
                want[i] = orc.retrieval_scores(sd, b, cfg, feats=f[i:i + 1].expand(nb, -1, -1)).numpy()
        got = host[:nb, :nb].astype(np.float64)
        oracle = {"block": nb, "max_abs_diff": float(np.abs(got - want).max()), "max_score": float(np.abs(want).max()),
                  "rel": float(np.abs(got - want).max() / np.abs(want).max())}
    pairs = a.images * a.reports
    print(json.dumps({"metric": "ITM retrieval scoring", "value": pairs / (ms_pairs / 1e3), "unit": "pairs/s", "n_gpus": world,
                      "images": a.images, "reports": a.reports, "pair_batch": a.pair_batch, "ms_pairs": ms_pairs,
                      "image_features_per_s": a.images / (ms_feat / 1e3), "ms_image_features": ms_feat, "wall_s": wall,
                      "encoder_tflops_dense_per_gpu": pairs * 8.1071e10 / (ms_pairs / 1e3) / 1e12 / world,
                      "sharding": "image rows over ranks, no data-path collective (all_gather of the result for the metrics)",
                      "hit@1/5/10": [res["i2t_retrieval"][k] for k in ("R@1", "R@5", "R@10")], "mrr": float(mrr), "oracle_sub_block": oracle,
                      "dtype": "bf16", "data": "synthetic (random-init weights: metrics are chance level by construction)"}))
if world > 1:
    torch.cuda.synchronize()
    torch.distributed.barrier()
    model._cxrbert._release_engine()
    torch.distributed.destroy_process_group()
