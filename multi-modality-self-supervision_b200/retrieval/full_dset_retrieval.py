"""Scoring path of /root/reference/Downstream_task/Retrieval/full_dset_retrieval.py on the B200 engine.

  data_processing           :186-218   report tokens -> (cls_tok, input_ids, attn_masks[L] 1-D, segment, sep_tok)
  test                      :461-510   model.eval(); softmax(logits)[:, 1] per (image, report) pair
  compute_ranks             :250-276   rank of the first aligned report inside each group of `eval_len_size`
  compute_recall_precision  :278-315   Recall@K / Precision@K, K in {1, 5, 10}
  compute_mrr, evaluate     :317-339   MRR, Hit@{1,5,10}

What changes: the reference forwards one pair at a time through ResNet-50 + BERT and keeps a Python
list of 0-d tensors; here ResNet-50 grid features are computed ONCE per image and cached in HBM (1 MB per image in
bf16), the joint encoder runs on batches of pairs through mv_forward (1-D padding mask == Bidirectional mode with the
sample's text length; no MLM head), and the match probabilities are written by mv_itm_match_prob straight into the
[images x reports] similarity matrix on the device — one D2H copy at the end.  Rank metrics are host integer work,
as in the reference (numpy; same argsort so ties resolve identically).
"""
import numpy as np
import torch

from .. import _lib

PAD, CLS, SEP = 0, 101, 102


def data_processing(encoded_sentence, seq_len, num_image_embeds, pad=PAD, cls=CLS, sep=SEP):
    """full_dset_retrieval.py:199-218 for an already tokenised + vocabulary-mapped report (<= seq_len ids)."""
    ids = list(encoded_sentence)[:seq_len] + [sep]
    t_len = len(ids)
    n_pad = seq_len + 1 - t_len
    attn = [1] * (num_image_embeds + 2) + [1] * t_len + [pad] * n_pad
    return dict(cls_tok=torch.tensor([cls]), input_ids=torch.tensor(ids + [pad] * n_pad), attn_masks=torch.tensor(attn),
                segment=torch.ones(seq_len + 1, dtype=torch.long), sep_tok=torch.tensor([sep]), t_len=t_len)


class RetrievalScorer:
    """Batched pair scorer over cached image features.

    scorer = RetrievalScorer(model)                      # CXRBertForRetrieval (or CXRBERT), on a CUDA device, .eval()
    feats  = scorer.image_features(images)               # [Ni, grid, 2048] device, activation dtype (cached by caller)
    sims   = scorer.score_matrix(feats, input_ids, t_len)  # [Ni, Nt] fp32 device: P(match | image i, report t)
    """

    def __init__(self, model, pair_batch=None):
        self.model = model
        self.owner = getattr(model, "_cxrbert", model)
        self.pair_batch = int(pair_batch or getattr(self.owner.args, "max_micro_batch", 64))
        self.eng = self.owner.engine(self.pair_batch)
        self.pair_batch = min(self.pair_batch, self.eng.max_batch)

    @torch.no_grad()
    def image_features(self, images, image_batch=32):
        """ResNet-50 grid features of every image, once (eval-mode BatchNorm when the model is in eval mode)."""
        out = []
        for s in range(0, images.shape[0], image_batch):
            x = images[s:s + image_batch].to(self.eng.device, non_blocking=True)
            out.append(self.owner.grid_features(x, self.eng).clone())
        return torch.cat(out, 0)

    def _regions(self, region_idx, grid):
        if region_idx is not None:
            return torch.as_tensor(region_idx, dtype=torch.long)
        return self.owner.enc.img_encoder.sample_regions(grid)       # models/image.py:63-65 (fresh draw per forward)

    @torch.no_grad()
    def score_pairs(self, feats, img_index, input_ids, t_len, out=None, region_idx=None):
        """P(match) for pairs (feats[img_index[p]], input_ids[p]); all tensors on the device.  Returns `out` [P] fp32."""
        eng, dev = self.eng, self.eng.device
        d = eng.dims
        P = int(input_ids.shape[0])
        if out is None:
            out = torch.empty(P, dtype=torch.float32, device=dev)
        input_ids = torch.as_tensor(input_ids).to(dev, torch.int64)
        t_len = torch.as_tensor(t_len).to(dev, torch.int32)
        img_index = torch.as_tensor(img_index).to(dev, torch.int64)
        cls = torch.full((self.pair_batch,), CLS, dtype=torch.int64, device=dev)
        sep = torch.full((self.pair_batch,), SEP, dtype=torch.int64, device=dev)
        seg = torch.ones(self.pair_batch, d.T, dtype=torch.int64, device=dev)
        mode = torch.full((self.pair_batch,), _lib.MODE_BIDIR, dtype=torch.uint8, device=dev)   # 1-D padding mask
        train = bool(self.model.training)
        for s in range(0, P, self.pair_batch):
            e = min(P, s + self.pair_batch)
            n = e - s
            f = feats.index_select(0, img_index[s:e])
            batch = eng.make_batch(cls_tok=cls[:n], input_ids=input_ids[s:e], segment=seg[:n], sep_tok=sep[:n], mode=mode[:n],
                                   t_len=t_len[s:e], region_idx=self._regions(region_idx, feats.shape[1]), feats=f, train=train)
            eng.forward(batch)
            _lib.check(_lib.lib().mv_itm_match_prob(eng._h, out[s:e].data_ptr(), n, _lib.stream_ptr(dev)), "mv_itm_match_prob")
        return out

    @torch.no_grad()
    def score_matrix(self, feats, input_ids, t_len, region_idx=None, rows=None):
        """[Ni, Nt] similarity matrix (row i = image i against every report).  `rows` restricts the images scored by
        this rank (data-parallel sharding by query: ranks own disjoint row ranges, no collective)."""
        dev = self.eng.device
        Ni, Nt = int(feats.shape[0]), int(input_ids.shape[0])
        rows = range(Ni) if rows is None else rows
        input_ids = torch.as_tensor(input_ids).to(dev, torch.int64)
        t_len = torch.as_tensor(t_len).to(dev, torch.int32)
        rows = torch.as_tensor(list(rows), dtype=torch.int64, device=dev)
        sims = torch.empty(rows.numel(), Nt, dtype=torch.float32, device=dev)
        # pairs are laid out row-major, so consecutive batches walk one image's reports and the matrix fills in place;
        # images are taken in groups so the replicated token ids stay small (<= ~64k pairs per group)
        group = max(1, 65536 // max(1, Nt))
        for g in range(0, rows.numel(), group):
            r = rows[g:g + group]
            self.score_pairs(feats, r.repeat_interleave(Nt), input_ids.repeat(r.numel(), 1), t_len.repeat(r.numel()),
                             out=sims[g:g + r.numel()].view(-1), region_idx=region_idx)
        return sims


def test(args, model, eval_dataset):
    """full_dset_retrieval.py:461-510 — same return value (results, labels, eval_losses, idx_lst) for a DataLoader that
    yields the reference's 8-tuples (cls_tok, input_txt, attn_mask, input_img, segment, sep_tok, label, idx)."""
    model.eval()
    results, labels, idx_lst, eval_losses = [], [], [], []
    for batch in eval_dataset:
        cls_tok, input_txt, attn_mask, input_img, segment, sep_tok, label, idx = batch
        with torch.no_grad():
            logits = model(cls_tok, input_txt, attn_mask, segment, input_img, sep_tok)
        lab = torch.as_tensor(label)
        labels.extend(lab.tolist())
        idx_lst.extend(torch.as_tensor(idx).tolist())
        eval_losses.append(float(torch.nn.functional.cross_entropy(logits, lab.to(logits.device))))
        results.extend(torch.softmax(logits, dim=1)[:, 1].cpu())
    return results, labels, eval_losses, idx_lst


def _grouped(args, results, labels, idx_lst=None):
    n = int(args.eval_len_size)
    sims = np.array([float(r) for r in results], dtype=np.float32).reshape(-1, n)
    labs = np.asarray(labels).reshape(-1, n)
    ids = None if idx_lst is None else np.asarray(idx_lst).reshape(-1, n)
    order = np.argsort(sims, axis=1)[:, ::-1]            # the reference's per-row `np.argsort(sim)[::-1]`
    return sims, labs, ids, order, n


def compute_ranks(args, results, labels, idx_lst):
    """-> (i2t_ranks, t2i_ranks, Aligned_lst): 0-based rank of the first aligned candidate per group, `eval_len_size`
    when the group has none (then the reference pairs it with the LAST candidate of the ordering)."""
    sims, labs, ids, order, n = _grouped(args, results, labels, idx_lst)
    sorted_lab = np.take_along_axis(labs, order, axis=1) == 1
    has = sorted_lab.any(axis=1)
    first = sorted_lab.argmax(axis=1)
    rank = np.where(has, first, n)
    hit = np.where(has, first, n - 1)
    aligned = [[ids[g, order[g, hit[g]]], int(rank[g])] for g in range(len(rank))]
    ranks = [int(r) for r in rank]
    if getattr(args, "i2t", False):
        return ranks, [], aligned
    if getattr(args, "t2i", False):
        return [], ranks, aligned
    return [], [], aligned


def compute_recall_precision(args, results, labels, idx_lst):
    sims, labs, _, order, n = _grouped(args, results, labels, idx_lst)
    sorted_lab = np.take_along_axis(labs, order, axis=1)
    total = sorted_lab.sum(axis=1)
    recall, precision = [], []
    for k in (1, 5, 10):
        top = sorted_lab[:, :k].sum(axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            recall.append(np.mean(top / total))
        precision.append(np.mean(top / k))
    key = "i2t" if getattr(args, "i2t", False) else "t2i"
    pack = lambda v: {"R@1": round(v[0], 3), "R@5": round(v[1], 3), "R@10": round(v[2], 3)}
    return {key + "_recall": pack(recall), key + "_precision": pack(precision)}


def compute_mrr(ranks):
    return np.mean(np.reciprocal(np.array(ranks, dtype=float) + 1))


def evaluate(args, test_results, test_labels, idx_lst):
    i2t, t2i, aligned = compute_ranks(args, test_results, test_labels, idx_lst)
    rp = compute_recall_precision(args, test_results, test_labels, idx_lst)
    ranks = i2t if getattr(args, "i2t", False) else t2i
    accs = [sum(r < k for r in ranks) / len(ranks) for k in (1, 5, 10)]
    key = "i2t_retrieval" if getattr(args, "i2t", False) else "t2i_retrieval"
    return {key: {"R@1": accs[0], "R@5": accs[1], "R@10": accs[2]}}, aligned, compute_mrr(ranks), rp
