"""Shared helpers for the parity tests: golden fixtures, oracle batches, engine construction."""
import json
import os

import numpy as np
import torch

from oracle import medvill_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    cfg = orc.Cfg(**json.loads(str(g["cfg"])))
    return g, cfg


def golden_batch(g, cfg):
    """Re-create the exact synthetic batch the fixture was generated from (seeded, machine independent)."""
    batch = orc.synthetic_batch(cfg, int(g["B"]), int(g["seed"]), mode=int(g["mode"]), mixed=bool(int(g["mixed"])),
                                s2s_prob=float(g["s2s_prob"]), bi_prob=float(g["bi_prob"]))
    assert np.array_equal(batch["input_ids"], g["input_ids"]) and np.array_equal(batch["txt_labels"], g["txt_labels"])
    assert np.array_equal(batch["region_idx"], g["region_idx"]) and np.array_equal(batch["mode"], g["modes"])
    return batch


def dims_from_cfg(cfg, dropout=0.0):
    import medvill_b200 as m

    return m.EngineDims(hidden=cfg.hidden, heads=cfg.heads, layers=cfg.layers, inter=cfg.inter, vocab=cfg.vocab,
                        max_pos=cfg.max_pos, type_vocab=cfg.type_vocab, num_image_embeds=cfg.num_image_embeds,
                        seq_len=cfg.seq_len, img_hidden=cfg.img_hidden, grid=cfg.grid, ln_eps=cfg.ln_eps,
                        head_ln_eps=cfg.head_ln_eps, dropout_p=dropout)


def oracle_feats(params, batch):
    """ResNet-50 grid features [B, grid, 2048] from the CPU oracle (models/image.py:56-58)."""
    with torch.no_grad():
        fmap = orc.resnet50_trunk(params, batch["image"])
        return torch.flatten(fmap, start_dim=2).transpose(1, 2).contiguous()


def summarize(t):
    t = t.detach().double().flatten().cpu()
    idx = (torch.arange(8, dtype=torch.long) * (t.numel() - 1)) // 7
    return np.concatenate([[float(t.sum()), float(t.abs().sum()), float(t.norm())], t[idx].numpy()])
