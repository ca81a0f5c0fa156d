"""Host-side sample construction for MedViLL pre-training — drop-in mirror of
/root/reference/data/dataset_origin.py (`CXRDataset`): ITM pair sampling, BERT-style MLM corruption, padding / labels /
segments and the five self-attention masks, all driven by Python's global `random` in the reference's call order so
that a given `random.seed` yields bit-identical integer tensors (verified against the real class in
oracle/make_golden.py; fixtures in tests/golden/dataset_seed1234.npz).

`__getitem__` returns the reference's 9-tuple
    (cls_tok, input_ids, txt_labels, attn_masks, image, segment, is_aligned, sep_tok, itm_prob).
With `args.compact_masks = True` the [L, L] int64 mask (1.5 MB / sample, 97 MB H2D per step at batch 64) is replaced by
the two integers `(mode, t_len)` the CUDA attention kernel consumes; the mask is then generated on the fly on the GPU.

Deviations from the shipped reference, both documented in SURVEY.md App. B: `self.disturbing_mask` (an
AttributeError at dataset_origin.py:104) reads `args.disturbing_mask`, and the Non-cross mode uses the standard
436-slot label layout (the shipped variant appends a 437th label while the mask stays 436 x 436).
"""
import json
import os
import random
import re

import torch
from torch.utils.data import Dataset

MODE_BIDIR, MODE_S2S, MODE_BAR, MODE_NONCROSS = 0, 1, 2, 3


def truncate_txt(txt_tokens, max_seq_len):
    del txt_tokens[max_seq_len:]


def _norm_label(s):
    return " ".join(sorted(re.sub(r"[^0-9a-z ]+", " ", str(s).lower()).split()))


def labels_equal(a, b):
    """`fuzz.token_sort_ratio(a, b) == 100` (dataset_origin.py:224) — fuzzywuzzy when installed, else the equivalent
    token-sorted normalised comparison."""
    try:
        from fuzzywuzzy import fuzz

        return fuzz.token_sort_ratio(a, b) == 100
    except ImportError:
        return _norm_label(a) == _norm_label(b)


def build_mask(mode, A, L, t_len):
    """[L, L] int64 0/1 mask of the given mode (closed form of dataset_origin.py:138-167)."""
    q = torch.arange(L).unsqueeze(1)
    k = torch.arange(L).unsqueeze(0)
    if mode == MODE_BIDIR:
        m = (k < A + t_len).expand(L, L)
    elif mode == MODE_S2S:
        m = (k < A) | ((q >= A) & (k <= q))
    elif mode == MODE_BAR:
        m = (q < A) | (k < A) | (k <= q)
    else:
        m = (q < A) == (k < A)
    return m.to(torch.long).contiguous()


class CXRDataset(Dataset):
    def __init__(self, data_path, tokenizer, transforms, args, vocab=None):
        self.args = args
        self.data_dir = os.path.dirname(data_path)
        with open(data_path) as f:
            self.data = [json.loads(line) for line in f if line.strip()]
        self.seq_len = args.seq_len
        self.max_seq_len = args.max_seq_len - args.num_image_embeds
        self.transforms = transforms
        self.total_len = self.seq_len + args.num_image_embeds + 3
        self.tokenizer = tokenizer
        if vocab is None:
            vocab = self._load_vocab(args)
        self.vocab_stoi = vocab
        self.vocab_len = len(vocab)
        self.unk = "<unk>" if args.bert_model == "albert-base-v2" else "[UNK]"
        self.pad = "<pad>" if args.bert_model == "albert-base-v2" else "[PAD]"

    @staticmethod
    def _load_vocab(args):
        from transformers import AutoTokenizer, BertTokenizer

        name = {"bert-small-scratch": "google/bert_uncased_L-4_H-512_A-8", "bert-base-scratch": "bert-base-uncased"}.get(
            args.bert_model, args.bert_model)
        tok = (AutoTokenizer if "/" in name and "google" not in name else BertTokenizer).from_pretrained(name)
        return tok.get_vocab() if hasattr(tok, "get_vocab") else tok.vocab

    def __len__(self):
        return len(self.data)

    # ---- ITM: aligned pair or a report with a different label set (dataset_origin.py:211-235) ----
    def get_random_line(self):
        row = self.data[random.randint(0, len(self.data) - 1)]
        return row["text"], row["label"]

    def random_pair_sampling(self, idx):
        rec = self.data[idx]
        _, _, label_key, txt_key, img_key = rec.keys()
        itm_prob = random.random()
        if itm_prob > 0.5:
            return rec[txt_key], rec[img_key], 1, itm_prob
        for _ in range(300):
            other_txt, other_label = self.get_random_line()
            if not labels_equal(rec[label_key], other_label):
                return other_txt, rec[img_key], 0, itm_prob
        return None

    # ---- MLM: 15 % of tokens; 80 % [MASK], 10 % random id, 10 % kept (dataset_origin.py:183-209) ----
    def random_word(self, tokens):
        labels = []
        mask_id = self.vocab_stoi["[MASK]"]
        for i, tok in enumerate(tokens):
            r = random.random()
            if r >= 0.15:
                labels.append(-100)
                continue
            r /= 0.15
            if r < 0.8:
                tokens[i] = mask_id
            elif r < 0.9:
                tokens[i] = random.randrange(self.vocab_len)
            labels.append(tok)
        if all(l == -100 for l in labels):
            labels[0] = tokens[0]
            tokens[0] = mask_id
        return tokens, labels

    def pick_mode(self):
        a = self.args
        if a.Mixed:
            assert (a.s2s_prob + a.bi_prob) == 1.0
            return random.choices([MODE_BIDIR, MODE_S2S], weights=[a.bi_prob, a.s2s_prob])[0]
        if a.BAR_attn:
            return MODE_BAR
        if a.disturbing_mask:
            return MODE_NONCROSS
        return MODE_BIDIR

    def __getitem__(self, idx):
        from PIL import Image

        a = self.args
        origin_txt, img_path, is_aligned, itm_prob = self.random_pair_sampling(idx)
        image = self.transforms(Image.open(os.path.join(self.data_dir, img_path)).convert("RGB"))
        words = self.tokenizer(origin_txt)
        truncate_txt(words, self.seq_len)
        ids = [self.vocab_stoi[w] if w in self.vocab_stoi else self.vocab_stoi[self.unk] for w in words]
        ids, lab = self.random_word(ids)
        ids.append(self.vocab_stoi["[SEP]"])
        lab.append(-100)
        t_len = len(ids)
        n_pad = self.seq_len + 1 - t_len
        ids += [self.vocab_stoi[self.pad]] * n_pad
        lab += [-100] * n_pad
        A = a.num_image_embeds + 2
        txt_labels = torch.tensor([-100] * A + lab)
        mode = self.pick_mode()
        if getattr(a, "compact_masks", False):
            attn = torch.tensor([mode, t_len], dtype=torch.long)
        elif mode == MODE_BIDIR and a.attn_1d and not a.Mixed:
            attn = torch.tensor([1] * (A + t_len) + [0] * n_pad)
        else:
            attn = build_mask(mode, A, self.total_len, t_len)
        return (torch.tensor([self.vocab_stoi["[CLS]"]]), torch.tensor(ids), txt_labels, attn, image,
                torch.ones(self.seq_len + 1, dtype=torch.long), torch.tensor(is_aligned), torch.tensor([self.vocab_stoi["[SEP]"]]),
                itm_prob)
