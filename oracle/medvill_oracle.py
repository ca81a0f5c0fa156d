"""TEST INFRASTRUCTURE — CPU oracle for the MedViLL pre-training step.  NOT product code.

A plain restatement (numpy for the integer work, torch-CPU tensor algebra for the floating-point work; no
transformers / torchvision modules) of the reference path named by BASELINE.json's north_star.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.

Parity status: PINNED against the reference's own modules.  oracle/make_golden.py imports the unmodified
/root/reference/models/{cxrbert_origin,image}.py and data/dataset_origin.py under the shims in oracle/ref_shim.py,
runs them on the same seeded inputs and asserts equality with this restatement before writing tests/golden/*.npz.
The third-party arithmetic under the reference (HF `transformers` BertModel, unpinned 3.x upstream; torchvision
ResNet-50) is therefore pinned to the versions installed here (transformers 5.5.0 eager path, torchvision 0.26).

Every function cites the reference lines it restates (paths relative to /root/reference).
"""
import math
import zlib
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

# attention-mask modes (README.md:27 names; data/dataset_origin.py:138-176)
MODE_BIDIR = 0      # full bidirectional / attn_1d   : k < A + t_len
MODE_S2S = 1        # Seq2Seq                        : k < A or (q >= A and A <= k <= q)
MODE_BAR = 2        # Bidirectional Auto-Regressive  : q < A or k < A or k <= q
MODE_NONCROSS = 3   # Non-cross ("disturbing_mask")  : (q < A) == (k < A)
# report-generation fine-tune variants (Downstream_task/report_generation_and_vqa/sc/data_loader.py:394-408): the causal
# block stops at the real text end and padded rows see the prefix only
MODE_S2S_FT = 4     # k < A or (A <= q < A + t_len and k <= q)
MODE_BAR_FT = 5     # q < A or k < A or (q < A + t_len and k <= q)

PAD, UNK, CLS, SEP, MASK = 0, 100, 101, 102, 103


@dataclass
class Cfg:
    hidden: int = 768
    heads: int = 12
    layers: int = 12
    inter: int = 3072
    vocab: int = 30522
    max_pos: int = 512
    type_vocab: int = 2
    num_image_embeds: int = 180     # main_origin.py:137
    seq_len: int = 253              # main_origin.py:130
    img_size: int = 512             # main_origin.py:138
    img_hidden: int = 2048          # main_origin.py:133
    ln_eps: float = 1e-12           # BertConfig.layer_norm_eps (upstream)
    head_ln_eps: float = 1e-5       # models/cxrbert_origin.py:212
    dropout: float = 0.0

    @property
    def A(self):  # [CLS] + image regions + [SEP]
        return self.num_image_embeds + 2

    @property
    def T(self):  # text slots incl. trailing [SEP] and pads
        return self.seq_len + 1

    @property
    def L(self):  # data/dataset_origin.py:37
        return self.seq_len + self.num_image_embeds + 3

    @property
    def grid(self):  # ResNet-50 stride 32
        return (self.img_size // 32) ** 2


TINY = dict(hidden=128, heads=2, layers=2, inter=256, vocab=1000, max_pos=64, num_image_embeds=9, seq_len=20,
            img_size=128)


# =====================================================================================================
# Integer / host side  (bit-exact targets)
# =====================================================================================================
def mask_allowed(mode, q, k, A, t_len):
    """Closed-form predicate of data/dataset_origin.py:138-176 (q, k broadcastable integer arrays)."""
    q = np.asarray(q)
    k = np.asarray(k)
    if mode == MODE_BIDIR:
        return (k < A + t_len) & (q >= 0)
    if mode == MODE_S2S:
        return (k < A) | ((q >= A) & (k >= A) & (k <= q))
    if mode == MODE_BAR:
        return (q < A) | (k < A) | (k <= q)
    if mode == MODE_NONCROSS:
        return (q < A) == (k < A)
    if mode == MODE_S2S_FT:
        return ((k < A) | ((q >= A) & (q < A + t_len) & (k <= q))) & (q >= 0)
    if mode == MODE_BAR_FT:
        return (q < A) | (k < A) | ((q < A + t_len) & (k <= q))
    raise ValueError("unknown mask mode %r" % (mode,))


def attention_mask(mode, A, L, t_len):
    """[L, L] int64 0/1 mask for one sample (what CXRDataset.__getitem__ returns as attn_masks_tensor)."""
    q = np.arange(L)[:, None]
    k = np.arange(L)[None, :]
    return mask_allowed(mode, q, k, A, t_len).astype(np.int64)


def dataset_mask_construction(mode, num_image_embeds, seq_len, n_input_ids_padded, t_len):
    """Step-by-step numpy restatement of the tensor construction at data/dataset_origin.py:138-167, kept separate
    from the closed form above so the two can be checked against each other (and against the real CXRDataset)."""
    A = num_image_embeds + 2
    L = seq_len + num_image_embeds + 3
    attn_masks = [1] * A + [1] * t_len + [0] * (seq_len + 1 - t_len)       # :113-127
    full_attn = np.tile(np.asarray(attn_masks, dtype=np.int64)[None, :], (L, 1))  # :138-139
    ext = np.zeros((L, L), dtype=np.int64)                                   # :141
    st, end = A, A + n_input_ids_padded                                      # :142 (len AFTER padding, :122)
    ext[:, :A] = 1                                                           # :143
    tril = np.tril(np.ones((L, L), dtype=np.int64))                          # :38
    ext[st:end, st:end] = tril[:end - st, :end - st]                         # :144-145
    if mode == MODE_BIDIR:
        return full_attn
    if mode == MODE_S2S:
        return ext
    if mode == MODE_BAR:
        ext = ext.copy()
        ext[:A, :] = 1                                                       # :160
        return ext
    if mode == MODE_NONCROSS:
        base = np.zeros((L, L), dtype=np.int64)                              # :164-166
        base[:A, :A] = 1
        base[A:, A:] = 1
        return base
    raise ValueError(mode)


def random_word(tokens, rng, vocab_len, mask_id=MASK):
    """data/dataset_origin.py:183-209 — BERT 15 % / 80-10-10 corruption; same MT19937 call order."""
    tokens = list(tokens)
    output_label = []
    for i, token in enumerate(tokens):
        prob = rng.random()
        if prob < 0.15:
            prob /= 0.15
            if prob < 0.8:
                tokens[i] = mask_id
            elif prob < 0.9:
                tokens[i] = rng.randrange(vocab_len)
            output_label.append(token)
        else:
            tokens[i] = token
            output_label.append(-100)
    if all(o == -100 for o in output_label):
        output_label[0] = tokens[0]
        tokens[0] = mask_id
    return tokens, output_label


def label_match(a, b):
    """Stand-in for fuzz.token_sort_ratio(a, b) == 100 (fuzzywuzzy is not installed)."""
    import re

    def norm(s):
        return " ".join(sorted(re.sub(r"[^0-9a-z ]+", " ", str(s).lower()).split()))

    return norm(a) == norm(b)


def random_pair_sampling(data, idx, rng):
    """data/dataset_origin.py:211-235.  data: list of dicts with keys (id, split, label, text, img) in that order."""
    rec = data[idx]
    _, _, label_k, txt_k, img_k = rec.keys()
    d_label, d_txt, d_img = rec[label_k], rec[txt_k], rec[img_k]
    itm_prob = rng.random()
    if itm_prob > 0.5:
        return d_txt, d_img, 1, itm_prob
    for _ in range(300):
        r = rng.randint(0, len(data) - 1)
        random_txt, random_label = data[r]["text"], data[r]["label"]
        if not label_match(d_label, random_label):
            return random_txt, d_img, 0, itm_prob
    return None


def build_sample(encoded_sentence, rng, cfg, vocab_len, mixed=False, bar=True, disturbing=False, attn_1d=False,
                 s2s_prob=1.0, bi_prob=0.0):
    """data/dataset_origin.py:102-181 minus image loading / tokenisation.  `encoded_sentence`: token ids already
    truncated to seq_len.  Returns a dict with the reference's per-sample tensors plus the compact
    (mode, t_len) description the CUDA path consumes."""
    N, S = cfg.num_image_embeds, cfg.seq_len
    A, T, L = cfg.A, cfg.T, cfg.L
    input_ids, txt_labels = random_word(encoded_sentence, rng, vocab_len)      # :102
    input_ids = input_ids + [SEP]                                              # :108
    txt_labels_t = txt_labels + [-100]                                         # :109
    t_len = len(input_ids)
    n_pad = S - len(input_ids) + 1                                             # :116-117
    input_ids = input_ids + [PAD] * n_pad                                      # :122
    txt_labels_t = txt_labels_t + [-100] * n_pad                               # :124
    labels = [-100] * A + txt_labels_t                                         # :110,126
    segment = [1] * T                                                          # :129
    if mixed:                                                                  # :152-155
        pick = rng.choices([MODE_BIDIR, MODE_S2S], weights=[bi_prob, s2s_prob])[0]
        mode = pick
    elif bar:                                                                  # :157-161
        mode = MODE_BAR
    elif disturbing:                                                           # :163-167
        mode = MODE_NONCROSS
    else:                                                                      # :169-176
        mode = MODE_BIDIR
    if mode == MODE_BIDIR and attn_1d and not mixed:
        mask = np.asarray([1] * A + [1] * t_len + [0] * (T - t_len), dtype=np.int64)
    else:
        mask = attention_mask(mode, A, L, t_len)
    return dict(cls_tok=np.asarray([CLS], dtype=np.int64), input_ids=np.asarray(input_ids, dtype=np.int64),
                txt_labels=np.asarray(labels, dtype=np.int64), attn_masks=mask,
                segment=np.asarray(segment, dtype=np.int64), sep_tok=np.asarray([SEP], dtype=np.int64),
                mode=mode, t_len=t_len)


def synthetic_batch(cfg, B, seed, mode=MODE_BAR, mixed=False, s2s_prob=1.0, bi_prob=0.0, min_len=None):
    """SURVEY.md §8(d) synthetic inputs: token ids U[999, vocab), t ~ U{min_len..seq_len}, Bernoulli(0.5) ITM labels,
    random_word corruption under random.Random(seed); image ~ N(0,1).  Integer tensors are numpy, image is torch."""
    import random

    rng = random.Random(seed)
    nrng = np.random.RandomState(seed)
    lo = min(999, cfg.vocab // 2)
    if min_len is None:
        min_len = min(16, cfg.seq_len)
    samples = []
    for _ in range(B):
        t = int(nrng.randint(min_len, cfg.seq_len + 1))
        toks = nrng.randint(lo, cfg.vocab, size=t).tolist()
        s = build_sample(toks, rng, cfg, cfg.vocab, mixed=mixed, bar=(mode == MODE_BAR and not mixed),
                         disturbing=(mode == MODE_NONCROSS), s2s_prob=s2s_prob, bi_prob=bi_prob)
        if not mixed and mode == MODE_S2S:
            s["mode"] = MODE_S2S
            s["attn_masks"] = attention_mask(MODE_S2S, cfg.A, cfg.L, s["t_len"])
        samples.append(s)
    g = torch.Generator().manual_seed(seed)
    batch = dict(
        cls_tok=np.stack([s["cls_tok"] for s in samples]), input_ids=np.stack([s["input_ids"] for s in samples]),
        txt_labels=np.stack([s["txt_labels"] for s in samples]), attn_masks=np.stack([s["attn_masks"] for s in samples]),
        segment=np.stack([s["segment"] for s in samples]), sep_tok=np.stack([s["sep_tok"] for s in samples]),
        is_aligned=nrng.randint(0, 2, size=B).astype(np.int64),
        mode=np.asarray([s["mode"] for s in samples], dtype=np.uint8),
        t_len=np.asarray([s["t_len"] for s in samples], dtype=np.int32),
        image=torch.randn(B, 3, cfg.img_size, cfg.img_size, generator=g),
        region_idx=np.sort(nrng.permutation(cfg.grid)[:cfg.num_image_embeds]).astype(np.int64),  # models/image.py:64-65
    )
    return batch


# =====================================================================================================
# Parameters (reference state_dict key names)
# =====================================================================================================
RESNET_LAYERS = ((4, 64, 3, 1), (5, 128, 4, 2), (6, 256, 6, 2), (7, 512, 3, 2))  # (Sequential idx, planes, blocks, stride)
RES = "enc.img_encoder.model."

ALIASES = {
    "enc.img_embeddings.token_type_embeddings.weight": "enc.txt_embeddings.token_type_embeddings.weight",
    "enc.img_embeddings.position_embeddings.weight": "enc.txt_embeddings.position_embeddings.weight",
    "enc.img_embeddings.LayerNorm.weight": "enc.txt_embeddings.LayerNorm.weight",
    "enc.img_embeddings.LayerNorm.bias": "enc.txt_embeddings.LayerNorm.bias",
    "mlm.predictions.decoder.weight": "enc.txt_embeddings.word_embeddings.weight",
}


def canonical_key(name):
    return ALIASES.get(name, name)


def resnet_param_shapes():
    """torchvision resnet50 children()[:-2] as nn.Sequential (models/image.py:50-52): names + shapes."""
    out = []

    def bn(prefix, c):
        out.extend([(prefix + ".weight", (c,)), (prefix + ".bias", (c,)), (prefix + ".running_mean", (c,)),
                    (prefix + ".running_var", (c,)), (prefix + ".num_batches_tracked", ())])

    out.append((RES + "0.weight", (64, 3, 7, 7)))
    bn(RES + "1", 64)
    inplanes = 64
    for idx, planes, blocks, stride in RESNET_LAYERS:
        for b in range(blocks):
            p = "%s%d.%d." % (RES, idx, b)
            out.append((p + "conv1.weight", (planes, inplanes, 1, 1)))
            bn(p + "bn1", planes)
            out.append((p + "conv2.weight", (planes, planes, 3, 3)))
            bn(p + "bn2", planes)
            out.append((p + "conv3.weight", (planes * 4, planes, 1, 1)))
            bn(p + "bn3", planes * 4)
            if b == 0:
                out.append((p + "downsample.0.weight", (planes * 4, inplanes, 1, 1)))
                bn(p + "downsample.1", planes * 4)
            inplanes = planes * 4
    return out


def bert_param_shapes(cfg):
    H, I, V = cfg.hidden, cfg.inter, cfg.vocab
    out = [
        ("enc.txt_embeddings.word_embeddings.weight", (V, H)),
        ("enc.txt_embeddings.position_embeddings.weight", (cfg.max_pos, H)),
        ("enc.txt_embeddings.token_type_embeddings.weight", (cfg.type_vocab, H)),
        ("enc.txt_embeddings.LayerNorm.weight", (H,)), ("enc.txt_embeddings.LayerNorm.bias", (H,)),
        ("enc.img_embeddings.img_embeddings.weight", (H, cfg.img_hidden)), ("enc.img_embeddings.img_embeddings.bias", (H,)),
    ]
    for l in range(cfg.layers):
        p = "enc.encoder.layer.%d." % l
        for n in ("query", "key", "value"):
            out += [(p + "attention.self.%s.weight" % n, (H, H)), (p + "attention.self.%s.bias" % n, (H,))]
        out += [(p + "attention.output.dense.weight", (H, H)), (p + "attention.output.dense.bias", (H,)),
                (p + "attention.output.LayerNorm.weight", (H,)), (p + "attention.output.LayerNorm.bias", (H,)),
                (p + "intermediate.dense.weight", (I, H)), (p + "intermediate.dense.bias", (I,)),
                (p + "output.dense.weight", (H, I)), (p + "output.dense.bias", (H,)),
                (p + "output.LayerNorm.weight", (H,)), (p + "output.LayerNorm.bias", (H,))]
    out += [
        ("enc.pooler.dense.weight", (H, H)), ("enc.pooler.dense.bias", (H,)),
        ("mlm.predictions.bias", (V,)),
        ("mlm.predictions.transform.dense.weight", (H, H)), ("mlm.predictions.transform.dense.bias", (H,)),
        ("mlm.predictions.transform.LayerNorm.weight", (H,)), ("mlm.predictions.transform.LayerNorm.bias", (H,)),
        ("itm.linear.weight", (2, H)), ("itm.linear.bias", (2,)),
    ]
    return out


def synth_params(cfg, seed=0, resnet=True):
    """Deterministic random-init weights keyed by the reference's state_dict names.  Not HF's initialiser: every
    tensor (biases, LayerNorm affine, BN affine included) is non-trivial so parity tests are discriminative."""
    params = {}

    def gen(name, shape, kind):
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        if kind == "norm_w":
            return 1.0 + 0.05 * torch.randn(shape, generator=g)
        if kind == "small":
            return 0.02 * torch.randn(shape, generator=g)
        if kind == "conv":
            fan_out = shape[0] * shape[2] * shape[3]
            return math.sqrt(2.0 / fan_out) * torch.randn(shape, generator=g)
        if kind == "img_proj":
            return (1.0 / math.sqrt(shape[1])) * torch.randn(shape, generator=g)
        raise ValueError(kind)

    for name, shape in bert_param_shapes(cfg):
        if name.endswith("LayerNorm.weight"):
            kind = "norm_w"
        elif name == "enc.img_embeddings.img_embeddings.weight":
            kind = "img_proj"
        else:
            kind = "small"
        params[name] = gen(name, shape, kind)
    if resnet:
        for name, shape in resnet_param_shapes():
            if name.endswith("num_batches_tracked"):
                params[name] = torch.zeros((), dtype=torch.long)
            elif name.endswith("running_mean"):
                params[name] = torch.zeros(shape)
            elif name.endswith("running_var"):
                params[name] = torch.ones(shape)
            elif len(shape) == 4:
                params[name] = gen(name, shape, "conv")
            elif name.endswith(".weight"):
                params[name] = gen(name, shape, "norm_w")
            else:
                params[name] = gen(name, shape, "small")
    return params


def trainable_names(cfg):
    """The 111.68 M-parameter trainable set: everything except the (entirely frozen) ResNet trunk
    (models/cxrbert_origin.py:66-70 — the 'unfreeze' loop iterates an empty list)."""
    return [n for n, _ in bert_param_shapes(cfg)]


# =====================================================================================================
# Floating-point forward  (torch-CPU tensor algebra)
# =====================================================================================================
def resnet50_trunk(params, x, bn_train=True, update_stats=False):
    """torchvision ResNet-50 v1.5 without avgpool/fc (models/image.py:50-56).  Train-mode BatchNorm (batch
    statistics) because the trainer calls model.train() (models/train_origin.py:72) on frozen weights."""

    def bn(prefix, t):
        rm, rv = params[prefix + ".running_mean"], params[prefix + ".running_var"]
        if not update_stats:
            rm, rv = rm.clone(), rv.clone()
        return F.batch_norm(t, rm, rv, params[prefix + ".weight"], params[prefix + ".bias"], training=bn_train,
                            momentum=0.1, eps=1e-5)

    x = F.conv2d(x, params[RES + "0.weight"], stride=2, padding=3)
    x = F.relu(bn(RES + "1", x))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for idx, planes, blocks, stride in RESNET_LAYERS:
        for b in range(blocks):
            p = "%s%d.%d." % (RES, idx, b)
            s = stride if b == 0 else 1
            identity = x
            out = F.relu(bn(p + "bn1", F.conv2d(x, params[p + "conv1.weight"])))
            out = F.relu(bn(p + "bn2", F.conv2d(out, params[p + "conv2.weight"], stride=s, padding=1)))
            out = bn(p + "bn3", F.conv2d(out, params[p + "conv3.weight"]))
            if b == 0:
                identity = bn(p + "downsample.1", F.conv2d(x, params[p + "downsample.0.weight"], stride=s))
            x = F.relu(out + identity)
    return x


def gelu(x):  # models/cxrbert_origin.py:176-181
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def tf_layer_norm(x, w, b, eps):  # models/cxrbert_origin.py:198-202
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return w * ((x - u) / torch.sqrt(s + eps)) + b


def encoder_layer(params, l, x, ext_mask, cfg, keep=None):
    """One post-LN BERT layer (upstream BertLayer; in-tree twin
    Downstream_task/report_generation_and_vqa/sc/pytorch_pretrained_bert/model.py:261-390), dropout = identity."""
    p = "enc.encoder.layer.%d." % l
    B, L, H = x.shape
    nh, d = cfg.heads, H // cfg.heads

    def lin(t, name):
        return t @ params[p + name + ".weight"].t() + params[p + name + ".bias"]

    def split(t):
        return t.view(B, L, nh, d).permute(0, 2, 1, 3)

    q, k, v = split(lin(x, "attention.self.query")), split(lin(x, "attention.self.key")), split(lin(x, "attention.self.value"))
    scores = (q @ k.transpose(-1, -2)) / math.sqrt(d) + ext_mask           # model.py:301-307
    probs = torch.softmax(scores, dim=-1)
    ctx = (probs @ v).permute(0, 2, 1, 3).reshape(B, L, H)
    a = F.layer_norm(lin(ctx, "attention.output.dense") + x, (H,), params[p + "attention.output.LayerNorm.weight"],
                     params[p + "attention.output.LayerNorm.bias"], cfg.ln_eps)
    h1 = lin(a, "intermediate.dense")
    out = F.layer_norm(lin(gelu(h1), "output.dense") + a, (H,), params[p + "output.LayerNorm.weight"],
                       params[p + "output.LayerNorm.bias"], cfg.ln_eps)
    if keep is not None:
        keep["l%d.q" % l], keep["l%d.k" % l], keep["l%d.v" % l] = q, k, v
        keep["l%d.ctx" % l], keep["l%d.attn_out" % l], keep["l%d.out" % l] = ctx, a, out
    return out


def extended_mask(attn_mask):
    """models/cxrbert_origin.py:75-85 — (1 - m.half()) * -10000, [B,1,L,L] or [B,1,1,L]."""
    m = torch.as_tensor(attn_mask)
    m = m[:, None, None, :] if m.dim() == 2 else m[:, None, :, :]
    return ((1.0 - m.to(torch.float16)) * -10000.0).to(torch.float32)


def joint_embeddings(params, batch, cfg, feats=None, keep=None):
    """models/cxrbert_origin.py:112-125 + 22-35 + upstream BertEmbeddings: the [B, L, H] encoder input."""
    H = cfg.hidden
    W = params["enc.txt_embeddings.word_embeddings.weight"]
    P = params["enc.txt_embeddings.position_embeddings.weight"]
    Ty = params["enc.txt_embeddings.token_type_embeddings.weight"]
    lw, lb = params["enc.txt_embeddings.LayerNorm.weight"], params["enc.txt_embeddings.LayerNorm.bias"]

    def ln(t):
        return F.layer_norm(t, (H,), lw, lb, cfg.ln_eps)

    cls_tok, sep_tok = torch.as_tensor(batch["cls_tok"]), torch.as_tensor(batch["sep_tok"])
    ids, seg = torch.as_tensor(batch["input_ids"]), torch.as_tensor(batch["segment"])
    ridx = torch.as_tensor(batch["region_idx"])
    B, T = ids.shape
    # upstream BertEmbeddings.word_embeddings = nn.Embedding(V, H, padding_idx=pad_token_id=0): the LOOKUP gradient of
    # id 0 ([PAD]) is dropped (the tied MLM-decoder gradient of row 0 is kept) — pinned by make_golden.py.
    emb = lambda i: F.embedding(i, W, padding_idx=PAD)
    cls_out = ln(emb(cls_tok) + P[:1][None] + Ty[0][None, None])             # :118 (position restarts at 0)
    sep_out = ln(emb(sep_tok) + P[:1][None] + Ty[0][None, None])             # :119
    if feats is None:
        fmap = resnet50_trunk(params, batch["image"])                        # image.py:56
        feats = torch.flatten(fmap, start_dim=2).transpose(1, 2).contiguous()  # image.py:57-58  [B, grid, 2048]
    if keep is not None:
        keep["feats"] = feats
    sampled = feats[:, ridx]                                                 # image.py:67
    img = sampled @ params["enc.img_embeddings.img_embeddings.weight"].t() + params["enc.img_embeddings.img_embeddings.bias"]
    img_out = ln(img + P[ridx][None] + Ty[0][None, None])                    # cxrbert_origin.py:24-33
    txt_out = ln(emb(ids) + P[:T][None] + Ty[seg])                           # :124
    return torch.cat([cls_out, img_out, sep_out, txt_out], 1)                # :125


def forward(params, batch, cfg, feats=None, keep=None):
    """CXRBERT.forward (models/cxrbert_origin.py:144-149): returns (mlm_logits [B,L,V], itm_logits [B,2])."""
    x = joint_embeddings(params, batch, cfg, feats=feats, keep=keep)
    if keep is not None:
        keep["emb"] = x
    ext = extended_mask(batch["attn_masks"])
    for l in range(cfg.layers):
        x = encoder_layer(params, l, x, ext, cfg, keep=keep)
    pooled = torch.tanh(x[:, 0] @ params["enc.pooler.dense.weight"].t() + params["enc.pooler.dense.bias"])  # :130
    t = x @ params["mlm.predictions.transform.dense.weight"].t() + params["mlm.predictions.transform.dense.bias"]
    t = tf_layer_norm(gelu(t), params["mlm.predictions.transform.LayerNorm.weight"],
                      params["mlm.predictions.transform.LayerNorm.bias"], cfg.head_ln_eps)            # :214-218
    logits = t @ params["enc.txt_embeddings.word_embeddings.weight"].t() + params["mlm.predictions.bias"]  # :235-238
    itm = pooled @ params["itm.linear.weight"].t() + params["itm.linear.bias"]                        # :172-173
    if keep is not None:
        keep["seq"], keep["pooled"], keep["mlm_transform"] = x, pooled, t
    return logits, itm


def losses(logits, itm, batch):
    """models/train_origin.py:62-63,118-126."""
    labels = torch.as_tensor(batch["txt_labels"])
    mlm = F.cross_entropy(logits.transpose(1, 2), labels, ignore_index=-100)
    itm_l = F.cross_entropy(itm, torch.as_tensor(batch["is_aligned"]))
    return mlm, itm_l, itm_l + mlm


def step_metrics(logits, itm, batch):
    """models/train_origin.py:133-146: ITM #correct, MLM #correct over labelled tokens, #labelled."""
    labels = torch.as_tensor(batch["txt_labels"])
    itm_correct = int(itm.argmax(-1).eq(torch.as_tensor(batch["is_aligned"])).sum())
    sel = labels != -100
    mlm_correct = int((logits.argmax(-1)[sel] == labels[sel]).sum())
    return itm_correct, mlm_correct, int(sel.sum())


def loss_and_grads(params, batch, cfg, feats=None, keep=None):
    """zero_grad -> backward of (itm + mlm) w.r.t. the trainable set (models/train_origin.py:129-130)."""
    names = trainable_names(cfg)
    leaf = dict(params)
    for n in names:
        leaf[n] = params[n].detach().clone().requires_grad_(True)
    logits, itm = forward(leaf, batch, cfg, feats=feats, keep=keep)
    mlm_l, itm_l, loss = losses(logits, itm, batch)
    loss.backward()
    grads = {n: (leaf[n].grad if leaf[n].grad is not None else torch.zeros_like(leaf[n])) for n in names}
    return dict(loss=loss.item(), mlm_loss=mlm_l.item(), itm_loss=itm_l.item(), logits=logits.detach(),
                itm_logits=itm.detach(), grads=grads)


def adamw_step(params, grads, state, lr, step, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0):
    """HF transformers-3.x AdamW.step with correct_bias=True (upstream; call site models/train_origin.py:60,131)."""
    b1, b2 = betas
    for n, g in grads.items():
        st = state.setdefault(n, dict(m=torch.zeros_like(params[n]), v=torch.zeros_like(params[n])))
        st["m"].mul_(b1).add_(g, alpha=1.0 - b1)
        st["v"].mul_(b2).addcmul_(g, g, value=1.0 - b2)
        denom = st["v"].sqrt().add_(eps)
        step_size = lr * math.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
        params[n] = params[n].addcdiv(st["m"], denom, value=-step_size)
        if weight_decay > 0.0:
            params[n] = params[n].add(params[n], alpha=-lr * weight_decay)
    return params


def attention_reference(q, k, v, mode, t_len, A, scale=None):
    """softmax(q k^T * scale + (1-m) * -10000) v for [B, nh, L, d] tensors with per-sample (mode, t_len)."""
    B, nh, L, d = q.shape
    scale = scale or 1.0 / math.sqrt(d)
    masks = np.stack([attention_mask(int(mode[b]), A, L, int(t_len[b])) for b in range(B)])
    ext = extended_mask(masks)
    probs = torch.softmax((q @ k.transpose(-1, -2)) * scale + ext, dim=-1)
    return probs @ v


# =====================================================================================================
# Label-conditioned retrieval scoring (BASELINE.json configs[4])
# =====================================================================================================
def retrieval_params(cfg, seed=0, bn_seed=7, itm_gain=60.0):
    """synth_params with (a) non-trivial BatchNorm running statistics — retrieval runs model.eval()
    (full_dset_retrieval.py:462), so BN reads them — and (b) the ITM weight scaled up so that match probabilities spread
    over (0, 1) instead of clustering at 0.5 (a random-init head is otherwise not discriminative for a ranking test)."""
    params = synth_params(cfg, seed=seed)
    g = torch.Generator().manual_seed(bn_seed)
    for k in sorted(params):
        if k.endswith("running_mean"):
            params[k] = 0.1 * torch.randn(params[k].shape, generator=g)
        elif k.endswith("running_var"):
            params[k] = 0.5 + torch.rand(params[k].shape, generator=g)
    params["itm.linear.weight"] = params["itm.linear.weight"] * itm_gain
    return params


def retrieval_pair(encoded_sentence, cfg):
    """CXR_Retrieval_Dataset.data_processing (Downstream_task/Retrieval/full_dset_retrieval.py:199-218):
    ids + [SEP], zero padding to seq_len + 1, 1-D attention mask over [CLS]+regions+[SEP] and the real text."""
    ids = list(encoded_sentence)[:cfg.seq_len] + [SEP]
    t_len = len(ids)
    n_pad = cfg.seq_len - t_len + 1
    return dict(cls_tok=np.asarray([CLS], dtype=np.int64), input_ids=np.asarray(ids + [PAD] * n_pad, dtype=np.int64),
                attn_masks=np.asarray([1] * (cfg.num_image_embeds + 2) + [1] * t_len + [PAD] * n_pad, dtype=np.int64),
                segment=np.ones(cfg.seq_len + 1, dtype=np.int64), sep_tok=np.asarray([SEP], dtype=np.int64), t_len=t_len)


def retrieval_logits(params, batch, cfg, feats=None):
    """CXRBertForRetrieval.forward (Downstream_task/Retrieval/retrieval.py:29-32): itm(pooled [CLS]) -> [B, 2]."""
    x = joint_embeddings(params, batch, cfg, feats=feats)
    ext = extended_mask(batch["attn_masks"])
    for l in range(cfg.layers):
        x = encoder_layer(params, l, x, ext, cfg)
    pooled = torch.tanh(x[:, 0] @ params["enc.pooler.dense.weight"].t() + params["enc.pooler.dense.bias"])
    return pooled @ params["itm.linear.weight"].t() + params["itm.linear.bias"]


def retrieval_scores(params, batch, cfg, feats=None):
    """test() of full_dset_retrieval.py:499-509: nn.Softmax(dim=1)(logits)[:, 1] — P(image and report match)."""
    return torch.softmax(retrieval_logits(params, batch, cfg, feats=feats), dim=1)[:, 1]


def retrieval_rank_metrics(sims, labels, idx_lst, group, direction="i2t"):
    """compute_ranks / compute_recall_precision / compute_mrr / evaluate (full_dset_retrieval.py:250-339) restated with
    explicit Python loops: per group of `group` candidates, order by descending similarity (np.argsort(sim)[::-1], so
    ties break as numpy's default sort breaks them), rank = position of the first aligned candidate (group size if none)."""
    sims = np.asarray(sims, dtype=np.float32).reshape(-1, group)
    labels = np.asarray(labels).reshape(-1, group)
    idx = np.asarray(idx_lst).reshape(-1, group)
    ranks, aligned = [], []
    per_k = {1: ([], []), 5: ([], []), 10: ([], [])}
    for g in range(sims.shape[0]):
        order = np.argsort(sims[g])[::-1]
        rank, at = group, order[-1]
        for pos, cand in enumerate(order):
            if labels[g][cand] == 1:
                rank, at = pos, cand
                break
        ranks.append(rank)
        aligned.append([idx[g][at], rank])
        ordered = [labels[g][c] for c in order]
        for k, (rec, prec) in per_k.items():
            top = np.array(ordered[:k]).sum()
            with np.errstate(divide="ignore", invalid="ignore"):
                rec.append(top / np.array(ordered).sum())
            prec.append(top / k)
    hits = {"R@%d" % k: sum(r < k for r in ranks) / len(ranks) for k in (1, 5, 10)}
    recall = {"R@%d" % k: round(np.mean(np.array(per_k[k][0])), 3) for k in (1, 5, 10)}
    precision = {"R@%d" % k: round(np.mean(np.array(per_k[k][1])), 3) for k in (1, 5, 10)}
    mrr = np.mean(np.reciprocal(np.array(ranks, dtype=float) + 1))
    return dict(ranks=ranks, aligned=aligned, hits=hits, recall=recall, precision=precision, mrr=mrr, direction=direction)


# =====================================================================================================
# Report-generation fine-tune step (BASELINE.json configs[4]).  Reference files are under
# Downstream_task/report_generation_and_vqa/sc/ : data_loader.py, pytorch_pretrained_bert/{model.py, optimization.py}, finetune.py
# =====================================================================================================
def finetune_cfg(**kw):
    """BERT-base fine-tune shapes: all 256 grid regions (pixel_full_sampling, model.py:36-54), max_len_b = 253 -> L = 512
    (finetune.py:68-76), LayerNorm eps 1e-5 everywhere (model.py:238,327,367)."""
    base = dict(num_image_embeds=256, seq_len=253, ln_eps=1e-5)
    base.update(kw)
    return Cfg(**base)


TINY_FT = dict(TINY, num_image_embeds=16, ln_eps=1e-5)       # 128x128 image -> 4x4 grid, every region used


def truncate_tokens_pair(tokens_a, tokens_b, max_len, rng, max_len_b=0, always_truncate_tail=False):
    """data_loader.py:24-59 with trunc_seg='b' (finetune.py:158): drop from the report until both fit; each removal
    draws rand() to pick head or tail unless always_truncate_tail."""
    while len(tokens_a) + len(tokens_b) > max_len:
        trunc = tokens_b          # max_len_a = 0, so either the max_len_b rule (:33-35) or trunc_seg == 'b' (:36-43) picks B
        if (not always_truncate_tail) and rng.random() < 0.5:
            del trunc[0]
        else:
            trunc.pop()


def preprocess4seq2seq(tokens_b, rng, cfg, mode="s2s", bar=False, max_pred=10, mask_prob=0.15, new_segment_ids=False,
                       always_truncate_tail=False):
    """Preprocess4Seq2seq.__call__ for tasks == 'report_generation' (data_loader.py:330-452) on token IDS.
    `rng` is a random.Random; the reference uses the module-level generator with the same call order:
    truncation draws, shuffle(cand_pos), one random() for the 50 % forced [SEP] mask.  Returns numpy int64 arrays."""
    N = cfg.num_image_embeds
    A = N + 2
    max_len = cfg.L
    tokens_a = [UNK] * N                                                    # :333
    tokens_b = list(tokens_b)
    truncate_tokens_pair(tokens_a, tokens_b, N + cfg.seq_len, rng, max_len_b=cfg.seq_len,
                         always_truncate_tail=always_truncate_tail)          # :336-338
    tokens = [CLS] + tokens_a + [SEP] + tokens_b + [SEP]                    # :340
    if new_segment_ids and mode == "s2s":                                   # :342-348
        segment_ids = [4] * A + [5] * (len(tokens_b) + 1)
    else:
        segment_ids = [0] * A + [1] * (len(tokens_b) + 1)
    n_pred = min(max_pred, max(1, int(round(len(tokens_b) * mask_prob))))   # :352-353
    cand_pos = [i for i, tk in enumerate(tokens) if i >= A and tk != CLS]   # :360-364
    rng.shuffle(cand_pos)                                                   # :367
    if rng.random() > 0.5:                                                  # :368-372  force-mask the final [SEP]
        masked_pos = cand_pos[:n_pred - 1]
        masked_pos.append(len(tokens) - 1)
    else:
        masked_pos = cand_pos[:n_pred]
    masked_ids = [tokens[p] for p in masked_pos]                            # :374
    for p in masked_pos:
        tokens[p] = MASK                                                    # :376-377
    masked_weights = [1] * len(masked_ids)                                  # :383
    n_pad = max_len - len(tokens)                                           # :390-392
    input_ids = tokens + [PAD] * n_pad
    segment_ids = segment_ids + [0] * n_pad
    t_len = len(tokens_b) + 1
    if bar:                                                                 # :398-402
        md = MODE_BAR_FT
    elif mode == "s2s":                                                     # :405-408
        md = MODE_S2S_FT
    else:                                                                   # :410-412 ('bi': 1-D mask expanded)
        md = MODE_BIDIR
    # step-by-step construction (kept independent of the closed form, asserted equal below)
    mask = np.zeros((max_len, max_len), dtype=np.int64)
    st, end = A, A + t_len
    tril = np.tril(np.ones((max_len, max_len), dtype=np.int64))
    if bar:
        mask[:, :A] = 1
        mask[:A, :] = 1
        mask[st:end, st:end] = tril[:end - st, :end - st]
    elif mode == "s2s":
        mask[:, :A] = 1
        mask[st:end, st:end] = tril[:end - st, :end - st]
    else:
        mask[:] = np.asarray([1] * len(tokens) + [0] * n_pad, dtype=np.int64)[None, :]
    assert np.array_equal(mask, attention_mask(md, A, max_len, t_len))
    n_real = len(masked_ids)                                                # :415-419 (pad to max_pred; n_pred may be < len)
    if max_pred > n_pred:
        pad = max_pred - n_pred
        masked_ids = masked_ids + [0] * pad
        masked_pos = masked_pos + [0] * pad
        masked_weights = masked_weights + [0] * pad
    i64 = lambda v: np.asarray(v, dtype=np.int64)
    return dict(input_ids=i64(input_ids), segment_ids=i64(segment_ids), input_mask=mask, masked_ids=i64(masked_ids),
                masked_pos=i64(masked_pos), masked_weights=i64(masked_weights), mode=md, t_len=t_len, n_real=n_real)


def finetune_batch(cfg, B, seed, mode="s2s", bar=False, new_segment_ids=False, min_len=1):
    """Seeded synthetic fine-tune batch: report ids U[999, vocab), length U{min_len..seq_len+6} (some get truncated)."""
    import random

    rng = random.Random(seed)
    nrng = np.random.RandomState(seed)
    lo = min(999, cfg.vocab // 2)
    samples = []
    for _ in range(B):
        t = int(nrng.randint(min_len, cfg.seq_len + 7))
        samples.append(preprocess4seq2seq(nrng.randint(lo, cfg.vocab, size=t).tolist(), rng, cfg, mode=mode, bar=bar,
                                          new_segment_ids=new_segment_ids))
    g = torch.Generator().manual_seed(seed)
    st = lambda k: np.stack([s[k] for s in samples])
    return dict(input_ids=st("input_ids"), segment_ids=st("segment_ids"), input_mask=st("input_mask"), masked_ids=st("masked_ids"),
                masked_pos=st("masked_pos"), masked_weights=st("masked_weights"),
                mode=np.asarray([s["mode"] for s in samples], dtype=np.uint8), t_len=np.asarray([s["t_len"] for s in samples], dtype=np.int32),
                image=torch.randn(B, 3, cfg.img_size, cfg.img_size, generator=g))


def finetune_embeddings(params, batch, cfg, feats=None):
    """BertForPreTrainingLossMask.forward up to the encoder input (model.py:970-975; ImageBertEmbeddings :864-900;
    vendored BertEmbeddings :223-260).  Differences from pre-training: every grid region is used (vis_pe = arange),
    the prefix [SEP] keeps position A-1, prefix token types come from segment_ids, no padding_idx on the word table."""
    H, N = cfg.hidden, cfg.num_image_embeds
    A = N + 2
    W = params["enc.txt_embeddings.word_embeddings.weight"]
    P = params["enc.txt_embeddings.position_embeddings.weight"]
    Ty = params["enc.txt_embeddings.token_type_embeddings.weight"]
    lw, lb = params["enc.txt_embeddings.LayerNorm.weight"], params["enc.txt_embeddings.LayerNorm.bias"]
    ids, seg = torch.as_tensor(batch["input_ids"]), torch.as_tensor(batch["segment_ids"])
    if feats is None:
        fmap = resnet50_trunk(params, batch["image"])                          # model.py:49 (train-mode BN: model.train())
        feats = torch.flatten(fmap, start_dim=2).transpose(1, 2).contiguous()    # :50-51
    img = feats @ params["enc.img_embeddings.img_embeddings.weight"].t() + params["enc.img_embeddings.img_embeddings.bias"]
    tok = torch.cat([W[ids[:, :1]], img, W[ids[:, A - 1:A]]], dim=1)           # :879-886
    pos = torch.cat([P[:1], P[torch.arange(N)], P[N + 1:A]], dim=0)[None]      # :888-892
    img_out = tf_layer_norm(tok + pos + Ty[seg[:, :A]], lw, lb, cfg.ln_eps)    # :894-898
    T = ids.shape[1] - A
    txt_out = tf_layer_norm(W[ids[:, A:]] + P[:T][None] + Ty[seg[:, A:]], lw, lb, cfg.ln_eps)   # :240-258
    return torch.cat([img_out, txt_out], 1)


def finetune_loss(params, batch, cfg, feats=None, keep=None, drop_worst_ratio=0.0):
    """masked-LM loss of the fine-tune step (model.py:968-1054): heads on the <= max_pred gathered rows,
    CE(reduction='none') * weights, summed per sample; the int(B * (1 - drop_worst_ratio)) samples with the SMALLEST loss
    are kept (:1006-1007; finetune.py:179,440 default ratio 0 keeps all) and their sum is divided by their weights + 1e-5."""
    x = finetune_embeddings(params, batch, cfg, feats=feats)
    ext = extended_mask(batch["input_mask"])
    for l in range(cfg.layers):
        x = encoder_layer(params, l, x, ext, cfg)
    pos = torch.as_tensor(batch["masked_pos"])
    rows = torch.gather(x, 1, pos[:, :, None].expand(-1, -1, x.shape[-1]))     # :992-993, 1040
    t = rows @ params["mlm.predictions.transform.dense.weight"].t() + params["mlm.predictions.transform.dense.bias"]
    t = tf_layer_norm(gelu(t), params["mlm.predictions.transform.LayerNorm.weight"],
                      params["mlm.predictions.transform.LayerNorm.bias"], cfg.head_ln_eps)
    logits = t @ params["enc.txt_embeddings.word_embeddings.weight"].t() + params["mlm.predictions.bias"]
    ce = F.cross_entropy(logits.transpose(1, 2).float(), torch.as_tensor(batch["masked_ids"]), reduction="none")   # :1047-1048
    w = torch.as_tensor(batch["masked_weights"]).to(ce.dtype)
    per_sample = (ce * w).sum(-1)                                              # :1004-1005
    kept, idx = torch.topk(per_sample, int(per_sample.shape[0] * (1 - drop_worst_ratio)), largest=False)     # :1007
    denom = w.sum(-1)[idx].sum() + 1e-5                                        # :1009
    if keep is not None:
        keep["seq"], keep["logits"], keep["ce"], keep["kept"] = x, logits, ce, idx
    return (kept / denom).sum()                                                # :1010


FT_NO_GRAD = ("enc.pooler.dense.weight", "enc.pooler.dense.bias", "itm.linear.weight", "itm.linear.bias")


def finetune_trainable_names(cfg):
    """parameters that receive a gradient in the fine-tune step: the pooled output feeds nothing (model.py:1041-1042
    discards seq_relationship), so the pooler has grad None and BertAdam skips it; there is no ITM head."""
    return [n for n in trainable_names(cfg) if n not in FT_NO_GRAD]


def finetune_loss_and_grads(params, batch, cfg, feats=None, keep=None, drop_worst_ratio=0.0):
    names = finetune_trainable_names(cfg)
    leaf = dict(params)
    for n in names:
        leaf[n] = params[n].detach().clone().requires_grad_(True)
    loss = finetune_loss(leaf, batch, cfg, feats=feats, keep=keep, drop_worst_ratio=drop_worst_ratio)
    loss.backward()
    grads = {n: (leaf[n].grad if leaf[n].grad is not None else torch.zeros_like(leaf[n])) for n in names}
    return dict(loss=loss.item(), grads=grads)


def warmup_linear(x, warmup=0.002):
    """optimization.py:45-48"""
    if x < warmup:
        return x / warmup
    return max((x - 1.0) / (warmup - 1.0), 0)


def bert_adam_decays(name):
    """finetune.py:383-389: no weight decay for names containing 'bias', 'LayerNorm.bias', 'LayerNorm.weight'"""
    return not any(nd in name for nd in ("bias", "LayerNorm.bias", "LayerNorm.weight"))


def bert_adam_step(params, grads, state, lr, step, t_total=-1, warmup=-1, b1=0.9, b2=0.999, e=1e-6, weight_decay=0.01,
                   max_grad_norm=1.0):
    """BertAdam.step (optimization.py:112-182): per-parameter clip_grad_norm_, no bias correction, decoupled decay
    added to the update, warmup_linear schedule evaluated at the parameter's own step counter (`step`, 0-based)."""
    lr_sched = lr * warmup_linear(step / t_total, warmup) if t_total != -1 else lr
    out = {}
    for n, g in grads.items():
        st = state.setdefault(n, dict(m=torch.zeros_like(params[n]), v=torch.zeros_like(params[n])))
        if max_grad_norm > 0:
            coef = max_grad_norm / (float(torch.linalg.vector_norm(g.float(), 2)) + 1e-6)
            g = g * min(coef, 1.0)
        st["m"].mul_(b1).add_(g, alpha=1 - b1)
        st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
        update = st["m"] / (st["v"].sqrt() + e)
        if weight_decay > 0.0 and bert_adam_decays(n):
            update = update + weight_decay * params[n]
        out[n] = params[n] - lr_sched * update
    return out
