"""CPU checks of the report-generation fine-tune path (BASELINE.json configs[4]): the oracle
restatement against fixtures generated from the real reference (oracle/make_golden_finetune.py), and the product's host-side
logic (Preprocess4Seq2seq, schedules, state_dict key set) — bit-exact for every integer output."""
import json
import random
import types

import numpy as np
import pytest
import torch

from oracle import medvill_oracle as orc
from tests.util import load_golden, summarize

STEP_FIXTURES = ["finetune_tiny_s2s", "finetune_tiny_bar", "finetune_tiny_bi", "finetune_tiny_s2s_newseg", "finetune_tiny_s2s_dropworst"]


@pytest.mark.parametrize("name", STEP_FIXTURES)
def test_oracle_reproduces_reference_finetune_step(name):
    g, cfg = load_golden(name)
    batch = orc.finetune_batch(cfg, int(g["B"]), int(g["seed"]), mode=str(g["mode_name"]), bar=bool(int(g["bar"])),
                               new_segment_ids=bool(int(g["new_segment_ids"])))
    for k in ("input_ids", "segment_ids", "masked_ids", "masked_pos", "masked_weights"):
        assert np.array_equal(batch[k], g[k]), k
    assert np.array_equal(batch["mode"], g["modes"]) and np.array_equal(batch["t_len"], g["t_len"])
    params = orc.synth_params(cfg, seed=0)
    ratio = float(g["drop_worst_ratio"]) if "drop_worst_ratio" in g else 0.0      # Luo's drop-worst, model.py:1003-1010
    keep = {}
    out = orc.finetune_loss_and_grads(params, batch, cfg, keep=keep, drop_worst_ratio=ratio)
    assert abs(out["loss"] - float(g["loss"])) < 2e-5 * abs(float(g["loss"]))
    if ratio > 0:
        assert np.array_equal(np.sort(keep["kept"].numpy()), g["kept_samples"]) and len(g["kept_samples"]) == int(int(g["B"]) * (1 - ratio))
    names = [str(n) for n in g["grad_names"]]
    assert names == sorted(orc.finetune_trainable_names(cfg))
    for i, n in enumerate(names):
        got, want = summarize(out["grads"][n]), g["grad_summary"][i]
        assert np.abs(got - want).max() <= 2e-3 * max(1e-6, np.abs(want[:3]).max()), n
    # three BertAdam steps on the reference's gradients (moments carry over; step 0 has a zero scheduled rate)
    cur, state = {n: params[n].clone() for n in names}, {}
    for k in range(3):
        cur = orc.bert_adam_step(cur, {n: out["grads"][n].clone() for n in names}, state, lr=float(g["adam_lr"]), step=k,
                                 t_total=int(g["adam_t_total"]), warmup=float(g["adam_warmup"]))
    for i, n in enumerate(names):
        got, want = summarize(cur[n] - params[n]), g["adam_summary"][i]
        assert np.abs(got - want).max() <= 5e-3 * np.abs(want[:3]).max() + 1e-9, n


def test_finetune_masks_closed_form_equals_construction():
    """S2S_FT / BAR_FT predicates == the tensor construction of data_loader.py:394-408 for every text length, and they
    coincide with the pre-training masks exactly when there is no padding"""
    cfg = orc.Cfg(**orc.TINY_FT)
    rng = random.Random(1)
    for t in range(1, cfg.seq_len + 1):
        toks = list(range(300, 300 + t))
        for mode, bar, md in (("s2s", False, orc.MODE_S2S_FT), ("s2s", True, orc.MODE_BAR_FT), ("bi", False, orc.MODE_BIDIR)):
            s = orc.preprocess4seq2seq(toks, rng, cfg, mode=mode, bar=bar)     # asserts construction == closed form inside
            assert s["mode"] == md and s["t_len"] == t + 1
    full = cfg.seq_len + 1
    assert np.array_equal(orc.attention_mask(orc.MODE_S2S_FT, cfg.A, cfg.L, full), orc.attention_mask(orc.MODE_S2S, cfg.A, cfg.L, full))
    assert np.array_equal(orc.attention_mask(orc.MODE_BAR_FT, cfg.A, cfg.L, full), orc.attention_mask(orc.MODE_BAR, cfg.A, cfg.L, full))
    assert not np.array_equal(orc.attention_mask(orc.MODE_S2S_FT, cfg.A, cfg.L, 3), orc.attention_mask(orc.MODE_S2S, cfg.A, cfg.L, 3))


def test_product_preprocess_bit_exact_with_reference_fixture():
    """medvill_b200.report_generation.Preprocess4Seq2seq under random.seed == the reference's outputs (fixture), for the
    s2s / bi / bar / new_segment_ids variants; the compact (mode, t_len) form describes the same mask"""
    import medvill_b200  # noqa: F401
    from medvill_b200.report_generation import Preprocess4Seq2seq

    g, cfg = load_golden("finetune_preprocess")
    words = ["[PAD]"] + ["w%d" % i for i in range(1, cfg.vocab)]
    for tok, i in (("[UNK]", orc.UNK), ("[CLS]", orc.CLS), ("[SEP]", orc.SEP), ("[MASK]", orc.MASK)):
        words[i] = tok
    stoi = {w: i for i, w in enumerate(words)}
    indexer = lambda toks: [stoi[t] for t in toks]
    args = types.SimpleNamespace(tasks="report_generation")
    img = torch.zeros(3, 4, 4)
    for vi in range(int(g["n_pre_variants"])):
        p = "pre%d_" % vi
        v = json.loads(str(g[p + "variant"]))
        kw = dict(new_segment_ids=v["new_segment_ids"], truncate_config={"max_len_b": cfg.seq_len, "trunc_seg": "b", "always_truncate_tail": False},
                  mode=v["mode"], len_vis_input=cfg.num_image_embeds, image_loader=lambda path: img)
        for compact in (False, True):
            pipe = Preprocess4Seq2seq(args, 10, 0.15, words, indexer, cfg.L, v["bar"], compact_mask=compact, **kw)
            random.seed(v["seed"])
            for i in range(g[p + "tokens"].shape[0]):
                ids = [int(t) for t in g[p + "tokens"][i] if t >= 0]
                out = pipe(("unused.png", [words[t] for t in ids], None, None, None))
                assert out[0] == g[p + "input_ids"][i].tolist() and out[1] == g[p + "segment_ids"][i].tolist()
                assert out[3] == g[p + "masked_ids"][i].tolist() and out[4] == g[p + "masked_pos"][i].tolist()
                assert out[5] == g[p + "masked_weights"][i].tolist()
                md, tl = int(g[p + "modes"][i]), int(g[p + "t_len"][i])
                if compact:
                    assert out[2].tolist() == [md, tl]
                else:
                    assert np.array_equal(out[2].numpy(), orc.attention_mask(md, cfg.A, cfg.L, tl))


def test_schedules_and_bert_adam_bookkeeping():
    from medvill_b200.report_generation import BertAdam, warmup_linear

    for x in (0.0, 0.001, 0.05, 0.1, 0.5, 0.999, 1.0, 1.5):
        assert warmup_linear(x, 0.1) == orc.warmup_linear(x, 0.1)
    p = torch.nn.Parameter(torch.zeros(4))
    opt = BertAdam([{"params": [p], "weight_decay": 0.01}, {"params": [], "weight_decay": 0.0}], lr=3e-5, warmup=0.1, t_total=100)
    assert opt.scheduled_lr() == 0.0 and opt.weight_decay == 0.01           # step counter is 0-based: first update has lr * 0
    opt.state["step"] = 5
    assert opt.scheduled_lr() == 3e-5 * 0.5
    with pytest.raises(Exception):
        opt.step()                                                          # no engine owns this parameter: loud failure


def test_finetune_model_state_dict_uses_the_reference_rename_rule():
    """keys == {rename(k) for k in pre-training keys} minus the ITM head (finetune.py:338-339 + strict=False load)"""
    import medvill_b200  # noqa: F401
    from medvill_b200.config import BertConfig
    from medvill_b200.report_generation import BertForPreTrainingLossMask, pretrain_to_finetune_key

    cfg = orc.Cfg(**orc.TINY_FT)
    bc = BertConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
                    intermediate_size=cfg.inter, max_position_embeddings=cfg.max_pos, type_vocab_size=2)
    args = types.SimpleNamespace(img_hidden_sz=cfg.img_hidden, hidden_size=cfg.hidden, img_postion=True, img_encoding="fully_use_cnn",
                                 len_vis_input=cfg.num_image_embeds, img_size=cfg.img_size, max_len_b=cfg.seq_len)
    m = BertForPreTrainingLossMask(bc, args, len_vis_input=cfg.num_image_embeds)
    keys = set(m.state_dict())
    params = orc.synth_params(cfg, seed=0)
    want = {pretrain_to_finetune_key(k) for k in params if not k.startswith("itm.")}
    assert want <= keys
    extra = keys - want
    assert all(k.startswith("img_embeddings.") or "decoder" in k for k in extra), sorted(extra)[:5]   # aliased shared modules
    assert m.L == cfg.L and m.A == cfg.A
    sd = {}
    for k in keys:
        src = [n for n in params if pretrain_to_finetune_key(n) == k]
        ck = src[0] if src else orc.canonical_key("enc." + k if not k.startswith("cls.") else k.replace("cls.", "mlm."))
        sd[k] = params[ck]
    m.load_state_dict(sd, strict=True)
    assert torch.equal(m.cls.predictions.decoder.weight.detach(), params["enc.txt_embeddings.word_embeddings.weight"])
