"""TEST INFRASTRUCTURE — pins oracle/medvill_oracle.py against the real reference and writes tests/golden/*.npz.

Run in the build container (needs /root/reference):   python oracle/make_golden.py [--full]
  * imports the UNMODIFIED reference modules under oracle/ref_shim.py,
  * loads the oracle's deterministic weights into the reference CXRBERT (load_state_dict, strict),
  * runs reference forward / CE losses / backward on the same synthetic batch (region indices injected through
    torch.randperm, dropout = 0) and asserts agreement with the restatement,
  * runs the reference CXRDataset (fake tokenizer + JSONL) under random.seed and asserts bit-equality with the
    oracle's random_pair_sampling / random_word / padding / masks for every mask mode,
  * commits small fixtures: losses, ITM logits, logits at labelled rows (subsampled), per-parameter gradient
    summaries, one-step AdamW summaries, and the dataset integer vectors.
"""
import argparse
import json
import os
import random
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import medvill_oracle as orc  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def ref_args(cfg, **over):
    a = types.SimpleNamespace(
        bert_model="bert-base-scratch", img_hidden_sz=cfg.img_hidden, embedding_size=cfg.hidden, hidden_size=cfg.hidden,
        dropout_prob=0.0, img_postion=True, img_encoder="random-pixel", num_image_embeds=cfg.num_image_embeds,
        img_size=cfg.img_size, disturbing_mask=False, vocab_size=cfg.vocab, seq_len=cfg.seq_len, max_seq_len=512,
        Mixed=False, BAR_attn=True, attn_1d=False, s2s_prob=1.0, bi_prob=0.0, img_channel=3)
    for k, v in over.items():
        setattr(a, k, v)
    return a


def build_reference_model(cfg, params):
    from transformers import BertConfig

    cxr, _ = ref_shim.load_reference_models()
    kw = dict(vocab_size=cfg.vocab, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
              intermediate_size=cfg.inter, max_position_embeddings=cfg.max_pos, type_vocab_size=cfg.type_vocab,
              hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, layer_norm_eps=cfg.ln_eps, hidden_act="gelu")
    ref_shim.set_bert_config(**kw)
    config = BertConfig(attn_implementation="eager", **kw)
    model = cxr.CXRBERT(config, ref_args(cfg))
    sd = model.state_dict()
    new = {}
    for k in sd:
        ck = orc.canonical_key(k)
        if ck in params:
            new[k] = params[ck].clone()
        elif k.endswith("position_ids") or k.endswith("token_type_ids"):
            new[k] = sd[k]
        else:
            raise KeyError("reference state_dict key %s has no oracle parameter" % k)
    missing = [k for k in params if k not in new]
    assert not missing, "oracle parameters unknown to the reference: %s" % missing[:5]
    model.load_state_dict(new, strict=True)
    model.train()  # models/train_origin.py:72 (BatchNorm uses batch statistics; dropout p = 0)
    return model


class _RandpermInject:
    """models/image.py:64 draws torch.randperm(grid)[:N]; return a permutation whose head is region_idx."""

    def __init__(self, region_idx, grid):
        rest = [i for i in range(grid) if i not in set(region_idx.tolist())]
        self.perm = torch.as_tensor(list(region_idx.tolist()) + rest, dtype=torch.long)

    def __enter__(self):
        self._orig = torch.randperm
        torch.randperm = lambda n, *a, **k: self.perm.clone()
        return self

    def __exit__(self, *exc):
        torch.randperm = self._orig


def run_reference_step(model, batch, cfg):
    t = lambda k: torch.as_tensor(batch[k])
    for p in model.parameters():
        p.grad = None
    with _RandpermInject(batch["region_idx"], cfg.grid):
        mlm_out, itm_out = model(t("cls_tok"), t("input_ids"), t("attn_masks"), t("segment"), batch["image"], t("sep_tok"))
    mlm_loss = torch.nn.CrossEntropyLoss(ignore_index=-100)(mlm_out.transpose(1, 2), t("txt_labels"))
    itm_loss = torch.nn.CrossEntropyLoss()(itm_out, t("is_aligned"))
    loss = itm_loss + mlm_loss
    loss.backward()
    grads = {}
    for n, p in model.named_parameters():
        if p.requires_grad:
            grads[orc.canonical_key(n)] = p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)
    return dict(loss=loss.item(), mlm_loss=mlm_loss.item(), itm_loss=itm_loss.item(), logits=mlm_out.detach(),
                itm_logits=itm_out.detach(), grads=grads)


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def summarize(t):
    t = t.detach().double().flatten()
    idx = (torch.arange(8, dtype=torch.long) * (t.numel() - 1)) // 7  # 8 probes (repeats for tiny tensors)
    return np.concatenate([[float(t.sum()), float(t.abs().sum()), float(t.norm())], t[idx].numpy()])


def pin_model(name, cfg, B, seed, mode, mixed=False, s2s_prob=1.0, bi_prob=0.0):
    print("[pin] %s: B=%d L=%d mode=%s" % (name, B, cfg.L, "mixed" if mixed else mode))
    params = orc.synth_params(cfg, seed=0)
    batch = orc.synthetic_batch(cfg, B, seed, mode=mode, mixed=mixed, s2s_prob=s2s_prob, bi_prob=bi_prob)
    model = build_reference_model(cfg, params)
    ref = run_reference_step(model, batch, cfg)
    keep = {}
    mine = orc.loss_and_grads(params, batch, cfg, keep=keep)
    # ---- pin: restatement == reference ----
    e_logits = rel_err(mine["logits"], ref["logits"])
    e_itm = rel_err(mine["itm_logits"], ref["itm_logits"])
    print("   loss ref=%.6f oracle=%.6f | rel err logits %.2e itm %.2e" % (ref["loss"], mine["loss"], e_logits, e_itm))
    assert abs(ref["loss"] - mine["loss"]) < 2e-5 * max(1.0, abs(ref["loss"])), "loss mismatch"
    assert e_logits < 5e-5 and e_itm < 5e-5, "logit mismatch"
    trainable = set(orc.trainable_names(cfg))
    assert set(ref["grads"].keys()) == trainable, "trainable set differs: %s" % (set(ref["grads"]) ^ trainable)
    worst = 0.0
    for n in sorted(trainable):
        # key-bias gradients are analytically zero (softmax is invariant to a per-query constant): absolute floor
        diff = float((mine["grads"][n].double() - ref["grads"][n].double()).abs().max())
        e = diff / (float(ref["grads"][n].abs().max()) + 1e-6)
        worst = max(worst, e)
        assert e < 2e-3, "grad mismatch %s: %.3e" % (n, e)
    print("   worst grad rel err %.2e over %d tensors" % (worst, len(trainable)))
    itm_c, mlm_c, n_lab = orc.step_metrics(ref["logits"], ref["itm_logits"], batch)
    # ---- one AdamW step on the reference's gradients (HF-3.x formula restated) ----
    p1 = orc.adamw_step({n: params[n].clone() for n in trainable}, ref["grads"], {}, lr=1e-5, step=1)
    # ---- fixture ----
    labels = batch["txt_labels"]
    rows = np.argwhere(labels != -100)
    lab_logits = ref["logits"][rows[:, 0], rows[:, 1]]                  # [n_lab, V]
    cols = np.unique(np.concatenate([np.linspace(0, cfg.vocab - 1, 64).astype(np.int64), labels[labels != -100]]))
    out = dict(
        cfg=json.dumps(cfg.__dict__), B=B, seed=seed, mode=int(mode), mixed=int(mixed), s2s_prob=s2s_prob, bi_prob=bi_prob,
        loss=ref["loss"], mlm_loss=ref["mlm_loss"], itm_loss=ref["itm_loss"], itm_logits=ref["itm_logits"].numpy(),
        lab_rows=rows, lab_cols=cols, lab_logits=lab_logits[:, cols].numpy(),
        lab_lse=torch.logsumexp(lab_logits.double(), -1).numpy(), lab_argmax=lab_logits.argmax(-1).numpy(),
        itm_correct=itm_c, mlm_correct=mlm_c, n_labelled=n_lab,
        seq_sample=keep["seq"][:, :: max(1, cfg.L // 16), :: max(1, cfg.hidden // 32)].detach().numpy(),
        emb_sample=keep["emb"][:, :: max(1, cfg.L // 16), :: max(1, cfg.hidden // 32)].detach().numpy(),
        feats_sample=keep["feats"][:, :: max(1, cfg.grid // 8), ::64].detach().numpy(),
        grad_names=np.asarray(sorted(trainable)),
        grad_summary=np.stack([summarize(ref["grads"][n]) for n in sorted(trainable)]),
        adamw_summary=np.stack([summarize(p1[n] - params[n]) for n in sorted(trainable)]),
        input_ids=batch["input_ids"], txt_labels=labels, t_len=batch["t_len"], modes=batch["mode"],
        is_aligned=batch["is_aligned"], region_idx=batch["region_idx"],
    )
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("   wrote %s.npz" % name)


class _FakeTok:
    def __init__(self, vocab_size):
        self.vocab = {str(i): i for i in range(vocab_size)}
        self.vocab.update({"[PAD]": 0, "[UNK]": 100, "[CLS]": 101, "[SEP]": 102, "[MASK]": 103})


def pin_dataset():
    """Reference CXRDataset.__getitem__ vs oracle, all mask modes, under the same random.seed."""
    from PIL import Image

    ds_mod = ref_shim.load_reference_dataset()
    cfg = orc.Cfg(num_image_embeds=180, seq_len=253)
    vocab = 30522
    ds_mod.BertTokenizer = types.SimpleNamespace(from_pretrained=lambda *a, **k: _FakeTok(vocab))
    tmp = tempfile.mkdtemp()
    Image.fromarray(np.zeros((32, 32, 3), dtype=np.uint8)).save(os.path.join(tmp, "x.png"))
    nrng = np.random.RandomState(7)
    recs = []
    for i in range(24):
        t = int(nrng.randint(5, 300))
        recs.append(dict(id=i, split="train", label="L%d" % (i % 5), text=" ".join(str(x) for x in nrng.randint(999, vocab, size=t)),
                         img="x.png"))
    path = os.path.join(tmp, "train.jsonl")
    with open(path, "w") as f:
        for r in recs:
            f.write(json.dumps(r) + "\n")
    tfm = lambda im: torch.zeros(3, 4, 4)
    tok = lambda s: s.split()
    fixtures = {}
    variants = {
        "bar": dict(BAR_attn=True), "bidir": dict(BAR_attn=False), "bidir1d": dict(BAR_attn=False, attn_1d=True),
        "s2s": dict(Mixed=True, s2s_prob=1.0, bi_prob=0.0), "mixed": dict(Mixed=True, s2s_prob=0.75, bi_prob=0.25),
        "noncross": dict(BAR_attn=False, disturbing_mask=True),
    }
    data = [json.loads(l) for l in open(path)]
    for vname, over in variants.items():
        args = ref_args(cfg, **over)
        ds = ds_mod.CXRDataset(path, tok, tfm, args)
        random.seed(1234)
        ref_items = [ds[i] for i in range(len(ds))]
        rng = random.Random(1234)
        ids_all, lab_all, mode_all, tlen_all, al_all = [], [], [], [], []
        for i, it in enumerate(ref_items):
            cls_tok, input_ids, txt_labels, attn, _img, segment, is_aligned, sep_tok, itm_prob = it
            txt, _img_path, al, prob = orc.random_pair_sampling(data, i, rng)
            enc = [int(w) for w in txt.split()][:cfg.seq_len]
            s = orc.build_sample(enc, rng, cfg, vocab, mixed=bool(over.get("Mixed")), bar=bool(over.get("BAR_attn", True)) and not over.get("Mixed"),
                                 disturbing=bool(over.get("disturbing_mask")), attn_1d=bool(over.get("attn_1d")),
                                 s2s_prob=over.get("s2s_prob", 1.0), bi_prob=over.get("bi_prob", 0.0))
            if vname == "noncross":
                # shipped Non-cross path: labels get one extra slot (dataset_origin.py:104-107);
                # the oracle implements the standard 436-slot layout, so compare everything but the label length.
                assert txt_labels.numel() == cfg.L + 1
                assert np.array_equal(txt_labels.numpy()[cfg.A + 1:], s["txt_labels"][cfg.A:])
            else:
                assert np.array_equal(txt_labels.numpy(), s["txt_labels"]), (vname, i)
            assert prob == itm_prob and al == int(is_aligned)
            assert np.array_equal(input_ids.numpy(), s["input_ids"]), (vname, i)
            assert np.array_equal(attn.numpy(), s["attn_masks"]), (vname, i, "mask")
            assert np.array_equal(segment.numpy(), s["segment"]) and int(cls_tok) == orc.CLS and int(sep_tok) == orc.SEP
            # step-by-step construction == closed form
            if attn.dim() == 2:
                assert np.array_equal(orc.dataset_mask_construction(s["mode"], cfg.num_image_embeds, cfg.seq_len, cfg.T, s["t_len"]),
                                      s["attn_masks"])
            ids_all.append(s["input_ids"]); lab_all.append(s["txt_labels"]); mode_all.append(s["mode"])
            tlen_all.append(s["t_len"]); al_all.append(al)
        print("[pin] dataset variant %-9s: %d samples bit-exact (modes %s)" % (vname, len(ref_items), sorted(set(mode_all))))
        fixtures[vname + "_input_ids"] = np.stack(ids_all)
        fixtures[vname + "_txt_labels"] = np.stack(lab_all)
        fixtures[vname + "_mode"] = np.asarray(mode_all, dtype=np.uint8)
        fixtures[vname + "_t_len"] = np.asarray(tlen_all, dtype=np.int32)
        fixtures[vname + "_is_aligned"] = np.asarray(al_all, dtype=np.int64)
    fixtures["records"] = np.asarray([json.dumps(r) for r in recs])
    np.savez_compressed(os.path.join(GOLDEN, "dataset_seed1234.npz"), **fixtures)
    print("   wrote dataset_seed1234.npz")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also pin BERT-base config 1 (B=2, L=436; ~1 min CPU)")
    a = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    pin_dataset()
    tiny = orc.Cfg(**orc.TINY)
    pin_model("tiny_bar", tiny, B=3, seed=11, mode=orc.MODE_BAR)
    pin_model("tiny_s2s", tiny, B=3, seed=12, mode=orc.MODE_S2S)
    pin_model("tiny_noncross", tiny, B=3, seed=13, mode=orc.MODE_NONCROSS)
    pin_model("tiny_bidir", tiny, B=3, seed=14, mode=orc.MODE_BIDIR)
    pin_model("tiny_mixed", tiny, B=4, seed=15, mode=orc.MODE_S2S, mixed=True, s2s_prob=0.75, bi_prob=0.25)
    if a.full:
        pin_model("config1_bar", orc.Cfg(), B=2, seed=123, mode=orc.MODE_BAR)


if __name__ == "__main__":
    main()
