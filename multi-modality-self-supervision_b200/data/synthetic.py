"""Synthetic MedViLL pre-training batches: the shapes, dtypes and integer semantics of
`CXRDataset.__getitem__` without MIMIC-CXR — uniform token ids in [999, V), text length t ~ U{16..S}, the reference's
random_word corruption (driven by Python's `random`), Bernoulli(0.5) ITM labels, N(0,1) "normalised CXR" pixels.
Used by bench.py, __graft_entry__.smoke() and the trainer tests; it is product-side code (no oracle import)."""
import random
import types

import numpy as np
import torch

from .dataset_origin import CXRDataset, MODE_BAR, MODE_BIDIR, MODE_S2S, build_mask

PAD, CLS, SEP, MASK = 0, 101, 102, 103


def _corruptor(vocab):
    stub = types.SimpleNamespace(vocab_stoi={"[MASK]": MASK}, vocab_len=vocab)
    return lambda toks: CXRDataset.random_word(stub, toks)


def synthetic_batch(B, vocab=30522, seq_len=253, num_image_embeds=180, img_size=512, seed=123, mode=MODE_BAR, mixed=False,
                    s2s_prob=1.0, bi_prob=0.0, full_masks=False, pin=False, image_dtype=torch.float32):
    """Returns a dict of CPU tensors in the reference's layout (+ compact mode/t_len)."""
    rnd_state = random.getstate()
    random.seed(seed)
    nrng = np.random.RandomState(seed)
    corrupt = _corruptor(vocab)
    A, T, L = num_image_embeds + 2, seq_len + 1, seq_len + num_image_embeds + 3
    lo = min(999, vocab // 2)
    ids = np.zeros((B, T), dtype=np.int64)
    labels = np.full((B, L), -100, dtype=np.int64)
    modes = np.zeros(B, dtype=np.uint8)
    t_len = np.zeros(B, dtype=np.int32)
    for b in range(B):
        t = int(nrng.randint(min(16, seq_len), seq_len + 1))
        toks, lab = corrupt(nrng.randint(lo, vocab, size=t).tolist())
        toks.append(SEP)
        lab.append(-100)
        ids[b, :len(toks)] = toks
        labels[b, A:A + len(lab)] = lab
        t_len[b] = len(toks)
        modes[b] = random.choices([MODE_BIDIR, MODE_S2S], weights=[bi_prob, s2s_prob])[0] if mixed else mode
    random.setstate(rnd_state)
    g = torch.Generator().manual_seed(seed)
    out = dict(
        cls_tok=torch.full((B, 1), CLS, dtype=torch.long), sep_tok=torch.full((B, 1), SEP, dtype=torch.long),
        input_ids=torch.from_numpy(ids), txt_labels=torch.from_numpy(labels), segment=torch.ones(B, T, dtype=torch.long),
        is_aligned=torch.from_numpy(nrng.randint(0, 2, size=B).astype(np.int64)),
        mode=torch.from_numpy(modes), t_len=torch.from_numpy(t_len),
        image=(torch.randint(0, 256, (B, 3, img_size, img_size), generator=g, dtype=torch.uint8) if image_dtype == torch.uint8
               else torch.randn(B, 3, img_size, img_size, generator=g).to(image_dtype)),
    )
    if full_masks:
        out["attn_masks"] = torch.stack([build_mask(int(modes[b]), A, L, int(t_len[b])) for b in range(B)])
    else:
        out["attn_masks"] = torch.stack([out["mode"].long(), out["t_len"].long()], dim=1)   # compact form
    if pin and torch.cuda.is_available():
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


def as_tuple(batch):
    """the reference's 9-tuple order (data/dataset_origin.py:181), batched"""
    return (batch["cls_tok"], batch["input_ids"], batch["txt_labels"], batch["attn_masks"], batch["image"], batch["segment"],
            batch["is_aligned"], batch["sep_tok"], torch.zeros(batch["input_ids"].shape[0]))
