"""TEST INFRASTRUCTURE — pins the report-generation fine-tune restatement (oracle/medvill_oracle.py: preprocess4seq2seq,
finetune_loss, bert_adam_step) against the real reference and writes tests/golden/finetune_*.npz.

Run in the build container (needs /root/reference):   python oracle/make_golden_finetune.py [--full]
Reference modules, imported UNMODIFIED from Downstream_task/report_generation_and_vqa/sc under stubs for what this
container lacks (torch._six, boto3/botocore, h5py) plus oracle/ref_shim.py's torchvision / .cuda() shims:
  * data_loader.Preprocess4Seq2seq (s2s, bi and bar variants) under random.seed -> bit-equality of input_ids, segment_ids,
    the [L, L] mask, masked_ids / masked_pos / masked_weights with the restatement on the same MT19937 stream;
  * pytorch_pretrained_bert.model.BertForPreTrainingLossMask (tasks='report_generation', img_encoding='fully_use_cnn'):
    loss and every parameter gradient of one step, dropout 0;
  * pytorch_pretrained_bert.optimization.BertAdam with finetune.py's parameter groups: two steps.
"""
import argparse
import collections.abc
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
from oracle.make_golden import rel_err, summarize  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
SC = os.path.join(ref_shim.REF_ROOT, "Downstream_task", "report_generation_and_vqa", "sc")


def load_reference_finetune():
    ref_shim._install_shims()
    six = types.ModuleType("torch._six")
    six.container_abcs = collections.abc
    six.string_classes = (str,)
    six.int_classes = (int,)
    six.inf = float("inf")
    sys.modules.setdefault("torch._six", six)
    torch._six = sys.modules["torch._six"]
    for name in ("boto3", "botocore", "botocore.exceptions", "h5py"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["botocore.exceptions"].ClientError = Exception
    if SC not in sys.path:
        sys.path.insert(0, SC)
    import data_loader  # noqa: E402  (reference module)
    from pytorch_pretrained_bert import model as ref_model  # noqa: E402
    from pytorch_pretrained_bert import optimization as ref_optim  # noqa: E402
    return data_loader, ref_model, ref_optim


def ft_key(name):
    """finetune.py:338-339: pre-training checkpoint keys -> fine-tune model keys"""
    return name.replace("enc.", "").replace("mlm.", "cls.")


def build_reference_model(ref_model, cfg, params, type_vocab):
    config = ref_model.BertConfig(cfg.vocab, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
                                  intermediate_size=cfg.inter, hidden_act="gelu", hidden_dropout_prob=0.0,
                                  attention_probs_dropout_prob=0.0, max_position_embeddings=cfg.max_pos, type_vocab_size=type_vocab)
    args = types.SimpleNamespace(img_hidden_sz=cfg.img_hidden, hidden_size=cfg.hidden, img_postion=True, img_encoding="fully_use_cnn",
                                 len_vis_input=cfg.num_image_embeds)
    model = ref_model.BertForPreTrainingLossMask(config, args, num_labels=2, len_vis_input=cfg.num_image_embeds,
                                                 tasks="report_generation")
    sd = model.state_dict()
    new = {}
    for k in sd:
        src = [n for n in params if ft_key(n) == k]
        if src:
            new[k] = params[src[0]].clone()
        else:
            assert k.endswith("num_batches_tracked") or k.startswith("img_embeddings.") or "decoder" in k, k
    # aliased modules (shared LayerNorm / tables, tied decoder) appear under several keys
    for k in sd:
        if k not in new:
            ck = orc.canonical_key("enc." + k if not k.startswith("cls.") else k.replace("cls.", "mlm."))
            new[k] = params[ck].clone() if ck in params else sd[k]
    model.load_state_dict(new, strict=True)
    model.train()                               # finetune.py:414
    model.img_embeddings.dropout.p = 0.0        # hard-coded nn.Dropout(0.1) (model.py:875): parity runs use dropout 0
    return model


def pin_preprocess(data_loader, cfg, out):
    """Preprocess4Seq2seq.__call__ under random.seed vs the restatement under random.Random(seed)."""
    from PIL import Image

    tmp = tempfile.mkdtemp()
    img_path = os.path.join(tmp, "x.png")
    Image.new("RGB", (cfg.img_size, cfg.img_size)).save(img_path)
    words = ["[PAD]"] + ["w%d" % i for i in range(1, cfg.vocab)]
    for tok, i in (("[UNK]", orc.UNK), ("[CLS]", orc.CLS), ("[SEP]", orc.SEP), ("[MASK]", orc.MASK)):
        words[i] = tok
    stoi = {w: i for i, w in enumerate(words)}
    indexer = lambda toks: [stoi[t] for t in toks]
    args = types.SimpleNamespace(tasks="report_generation")
    nrng = np.random.RandomState(77)
    n_checked = 0
    for vi, (mode, bar, newseg) in enumerate((("s2s", False, False), ("bi", False, False), ("s2s", True, False), ("s2s", False, True))):
        pipe = data_loader.Preprocess4Seq2seq(args, 10, 0.15, words, indexer, cfg.L, bar, new_segment_ids=newseg,
                                              truncate_config={"max_len_b": cfg.seq_len, "trunc_seg": "b", "always_truncate_tail": False},
                                              mode=mode, len_vis_input=cfg.num_image_embeds)
        seed = 4000 + vi
        random.seed(seed)
        rng = random.Random(seed)
        recs = []
        for _ in range(12):
            t = int(nrng.randint(1, cfg.seq_len + 8))
            ids = nrng.randint(200, cfg.vocab, size=t).tolist()
            ref = pipe((img_path, [words[i] for i in ids], None, None, None))
            mine = orc.preprocess4seq2seq(ids, rng, cfg, mode=mode, bar=bar, new_segment_ids=newseg)
            r_ids, r_seg, r_mask, r_mids, r_mpos, r_mw = ref[0], ref[1], ref[2].numpy(), ref[3], ref[4], ref[5]
            assert r_ids == mine["input_ids"].tolist() and r_seg == mine["segment_ids"].tolist(), (mode, bar)
            assert np.array_equal(r_mask, mine["input_mask"]), (mode, bar)
            assert r_mids == mine["masked_ids"].tolist() and r_mpos == mine["masked_pos"].tolist(), (mode, bar)
            assert r_mw == mine["masked_weights"].tolist(), (mode, bar)
            recs.append(dict(tokens=ids, **{k: mine[k] for k in ("input_ids", "segment_ids", "masked_ids", "masked_pos", "masked_weights")},
                             mode=mine["mode"], t_len=mine["t_len"]))
            n_checked += 1
        p = "pre%d_" % vi
        out[p + "variant"] = json.dumps(dict(mode=mode, bar=bar, new_segment_ids=newseg, seed=seed))
        out[p + "tokens"] = np.asarray([r["tokens"] + [-1] * (cfg.seq_len + 8 - len(r["tokens"])) for r in recs])
        for k in ("input_ids", "segment_ids", "masked_ids", "masked_pos", "masked_weights"):
            out[p + k] = np.stack([r[k] for r in recs])
        out[p + "modes"] = np.asarray([r["mode"] for r in recs])
        out[p + "t_len"] = np.asarray([r["t_len"] for r in recs])
    out["n_pre_variants"] = 4
    print("[pin] Preprocess4Seq2seq: %d samples over 4 variants bit-exact (ids, segments, [L,L] masks, masked ids/pos/weights)" % n_checked)


def pin_step(name, ref_model, ref_optim, cfg, B, seed, mode="s2s", bar=False, new_segment_ids=False, drop_worst_ratio=0):
    type_vocab = 6 if new_segment_ids else 2
    cfg = orc.Cfg(**dict(cfg.__dict__, type_vocab=type_vocab))
    print("[pin] %s: B=%d L=%d mode=%s bar=%s new_segment_ids=%s drop_worst_ratio=%s" % (name, B, cfg.L, mode, bar, new_segment_ids,
                                                                                       drop_worst_ratio))
    params = orc.synth_params(cfg, seed=0)
    batch = orc.finetune_batch(cfg, B, seed, mode=mode, bar=bar, new_segment_ids=new_segment_ids)
    model = build_reference_model(ref_model, cfg, params, type_vocab)
    t = lambda k: torch.as_tensor(batch[k])
    for p in model.parameters():
        p.grad = None
    loss, _ = model(batch["image"], None, t("input_ids"), t("segment_ids"), t("input_mask"), t("masked_ids"), None,
                    masked_pos=t("masked_pos"), masked_weights=t("masked_weights"), task_idx=None, drop_worst_ratio=drop_worst_ratio)
    loss = loss.mean()                                                       # finetune.py:447
    loss.backward()
    inv = {ft_key(n): n for n in orc.trainable_names(cfg)}
    ref_grads, no_grad = {}, []
    for n, p in model.named_parameters():
        if n.startswith("img_encoder."):
            assert p.grad is None and not p.requires_grad, n
            continue
        if p.grad is None:
            no_grad.append(inv[n])
        else:
            ref_grads[inv[n]] = p.grad.detach().clone()
    assert sorted(no_grad) == sorted(n for n in orc.FT_NO_GRAD if n.startswith("enc.pooler")), no_grad
    keep = {}
    mine = orc.finetune_loss_and_grads(params, batch, cfg, keep=keep, drop_worst_ratio=drop_worst_ratio)
    print("   loss ref=%.6f oracle=%.6f" % (float(loss), mine["loss"]))
    assert abs(float(loss) - mine["loss"]) < 2e-5 * max(1.0, abs(float(loss)))
    assert set(ref_grads) == set(orc.finetune_trainable_names(cfg))
    worst = 0.0
    for n in sorted(ref_grads):
        diff = float((mine["grads"][n].double() - ref_grads[n].double()).abs().max())
        e = diff / (float(ref_grads[n].abs().max()) + 1e-6)
        worst = max(worst, e)
        assert e < 2e-3, "grad mismatch %s: %.3e" % (n, e)
    print("   worst grad rel err %.2e over %d tensors" % (worst, len(ref_grads)))
    # ---- BertAdam, finetune.py:383-395 groups; two steps on the same gradients (moments carry over) ----
    named = list(model.named_parameters())
    no_decay = ["bias", "LayerNorm.bias", "LayerNorm.weight"]
    groups = [{"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": 0.01},
              {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
    lr, t_total, warm = 3e-5, 100, 0.1
    opt = ref_optim.BertAdam(groups, lr=lr, warmup=warm, schedule="warmup_linear", t_total=t_total)
    grads0 = {k: v.clone() for k, v in ref_grads.items()}
    cur = {n: params[n].clone() for n in ref_grads}
    state = {}
    # the scheduled rate of the reference's step k uses state['step'] = k (0-based): 0 at the very first step, so pin
    # steps 1 and 2 after a throw-away step 0 whose update is lr * 0
    for k in range(3):
        for n, p in model.named_parameters():
            if inv.get(n) in grads0:
                p.grad = grads0[inv[n]].clone()
        opt.step()
        cur = orc.bert_adam_step(cur, {n: grads0[n].clone() for n in grads0}, state, lr=lr, step=k, t_total=t_total, warmup=warm)
        worst_p = 0.0
        for n, p in model.named_parameters():
            if inv.get(n) in cur:
                worst_p = max(worst_p, float((p.detach() - cur[inv[n]]).abs().max()))
        print("   BertAdam step %d: max |ref - oracle| = %.2e" % (k, worst_p))
        assert worst_p < 1e-7
    names = sorted(ref_grads)
    rows = np.argwhere(batch["masked_weights"] > 0)
    out = dict(cfg=json.dumps(cfg.__dict__), B=B, seed=seed, mode_name=mode, bar=int(bar), new_segment_ids=int(new_segment_ids),
               loss=float(loss), drop_worst_ratio=float(drop_worst_ratio), kept_samples=np.sort(keep["kept"].numpy()), input_ids=batch["input_ids"], segment_ids=batch["segment_ids"], masked_ids=batch["masked_ids"],
               masked_pos=batch["masked_pos"], masked_weights=batch["masked_weights"], modes=batch["mode"], t_len=batch["t_len"],
               ce=keep["ce"].detach().numpy(), lab_rows=rows,
               lab_lse=torch.logsumexp(keep["logits"].detach().double(), -1).numpy()[rows[:, 0], rows[:, 1]],
               seq_sample=keep["seq"][:, :: max(1, cfg.L // 16), :: max(1, cfg.hidden // 32)].detach().numpy(),
               grad_names=np.asarray(names), grad_summary=np.stack([summarize(ref_grads[n]) for n in names]),
               adam_lr=lr, adam_t_total=t_total, adam_warmup=warm,
               adam_summary=np.stack([summarize(cur[n] - params[n]) for n in names]))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("   wrote %s.npz" % name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also pin BERT-base (B=2, L=512; ~2 min CPU)")
    a = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    data_loader, ref_model, ref_optim = load_reference_finetune()
    tiny = orc.Cfg(**orc.TINY_FT)
    pre = dict(cfg=json.dumps(tiny.__dict__))
    pin_preprocess(data_loader, tiny, pre)
    np.savez_compressed(os.path.join(GOLDEN, "finetune_preprocess.npz"), **pre)
    pin_step("finetune_tiny_s2s", ref_model, ref_optim, tiny, B=4, seed=21)
    pin_step("finetune_tiny_bar", ref_model, ref_optim, tiny, B=3, seed=22, bar=True)
    pin_step("finetune_tiny_bi", ref_model, ref_optim, tiny, B=3, seed=23, mode="bi", new_segment_ids=False)
    pin_step("finetune_tiny_s2s_newseg", ref_model, ref_optim, tiny, B=3, seed=24, new_segment_ids=True)
    # Luo's drop-worst (model.py:1006-1010): int(5 * 0.7) = 3 of 5 samples kept
    pin_step("finetune_tiny_s2s_dropworst", ref_model, ref_optim, tiny, B=5, seed=26, drop_worst_ratio=0.3)
    if a.full:
        pin_step("finetune_base_s2s", ref_model, ref_optim, orc.finetune_cfg(), B=2, seed=25)


if __name__ == "__main__":
    main()
