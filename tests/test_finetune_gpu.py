"""GPU parity of the report-generation fine-tune step (BASELINE.json configs[4]) through the drop-in
BertForPreTrainingLossMask / BertAdam surface -> mv_forward / mv_backward / mv_bert_adam_step, against fixtures generated
from the real reference (tests/golden/finetune_*.npz) and the CPU oracle.

Tolerances: fp32 check mode — loss 1e-4 relative, gradient norms 3e-3, BertAdam parameter deltas 2e-2 of the tensor's delta
norm; bf16 — loss 1e-2 relative, per-tensor gradient cosine vs the oracle.  Masks (fine-tune S2S / BAR variants) bit-exact.
"""
import types

import numpy as np
import pytest
import torch

from oracle import medvill_oracle as orc
from tests.util import load_golden, summarize

pytestmark = pytest.mark.gpu

TINY = ["finetune_tiny_s2s", "finetune_tiny_bar", "finetune_tiny_bi", "finetune_tiny_s2s_newseg"]


def make_model(cfg, precision, params, max_batch=8):
    import medvill_b200  # noqa: F401
    from medvill_b200.config import BertConfig
    from medvill_b200.report_generation import BertForPreTrainingLossMask, pretrain_to_finetune_key

    bc = BertConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
                    intermediate_size=cfg.inter, max_position_embeddings=cfg.max_pos, type_vocab_size=cfg.type_vocab,
                    hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    args = types.SimpleNamespace(img_hidden_sz=cfg.img_hidden, hidden_size=cfg.hidden, img_postion=True, img_encoding="fully_use_cnn",
                                 len_vis_input=cfg.num_image_embeds, img_size=cfg.img_size, max_len_b=cfg.seq_len, precision=precision,
                                 max_micro_batch=max_batch)
    model = BertForPreTrainingLossMask(bc, args, len_vis_input=cfg.num_image_embeds)
    sd = {}
    for k in model.state_dict():
        src = [n for n in params if pretrain_to_finetune_key(n) == k]
        ck = src[0] if src else orc.canonical_key("enc." + k if not k.startswith("cls.") else k.replace("cls.", "mlm."))
        sd[k] = params[ck]
    model.load_state_dict(sd, strict=True)
    return model.to("cuda:0").train()


def fixture_batch(g, cfg):
    batch = orc.finetune_batch(cfg, int(g["B"]), int(g["seed"]), mode=str(g["mode_name"]), bar=bool(int(g["bar"])),
                               new_segment_ids=bool(int(g["new_segment_ids"])))
    assert np.array_equal(batch["input_ids"], g["input_ids"]) and np.array_equal(batch["masked_pos"], g["masked_pos"])
    return batch


def oracle_feats(params, batch):
    with torch.no_grad():
        return torch.flatten(orc.resnet50_trunk(params, batch["image"]), start_dim=2).transpose(1, 2).contiguous()


def run(name, precision, explicit_mask, drop_worst_ratio=0.0):
    g, cfg = load_golden(name)
    batch = fixture_batch(g, cfg)
    params = orc.synth_params(cfg, seed=0)
    model = make_model(cfg, precision, params)
    t = lambda k: torch.as_tensor(batch[k])
    feats = oracle_feats(params, batch)         # the trunk is gated separately (tests/test_model_gpu.py); isolate the step
    kw = dict(feats=feats) if explicit_mask else dict(feats=feats, mode=t("mode"), t_len=t("t_len"))
    out = model.finetune_step(None, t("input_ids"), t("segment_ids"), t("input_mask") if explicit_mask else None, t("masked_ids"),
                              t("masked_pos"), t("masked_weights"), optimizer=None, drop_worst_ratio=drop_worst_ratio, **kw)
    torch.cuda.synchronize()
    return g, cfg, batch, params, model, out


@pytest.mark.parametrize("name", TINY)
def test_finetune_step_fp32_matches_reference(name):
    from medvill_b200.report_generation import BertAdam

    g, cfg, batch, params, model, out = run(name, "fp32", explicit_mask=True)     # [B, L, L] masks -> classified on the device
    assert abs(out["loss"] - float(g["loss"])) <= 1e-4 * abs(float(g["loss"])), (out["loss"], float(g["loss"]))
    assert out["n_masked"] == pytest.approx(float(batch["masked_weights"].sum()), abs=1e-3)
    eng = model.engine()
    names = [str(n) for n in g["grad_names"]]
    scale = max(r[2] for r in g["grad_summary"])
    for i, n in enumerate(names):
        got, ref = summarize(eng.view(n, eng.grads)), g["grad_summary"][i]
        if n.endswith("attention.self.key.bias"):
            assert got[2] <= 1e-4 * scale
            continue
        assert abs(got[2] - ref[2]) <= 3e-3 * max(ref[2], 1e-6), "grad norm %s: got %.6e ref %.6e" % (n, got[2], ref[2])
    for n in orc.FT_NO_GRAD:                                           # pooler / ITM head: no gradient on this path
        assert float(eng.view(n, eng.grads).abs().max()) == 0.0, n
    # three BertAdam steps on the same gradients, as pinned in the fixture (step 0 has a zero scheduled rate)
    before = {n: eng.view(n).clone() for n in eng.pmap}
    grads = eng.grads.clone()
    opt = BertAdam([{"params": [p for p in model.parameters()], "weight_decay": 0.01}], lr=float(g["adam_lr"]),
                   warmup=float(g["adam_warmup"]), t_total=int(g["adam_t_total"]))
    for k in range(3):
        eng.grads.copy_(grads)
        opt.step()
    torch.cuda.synchronize()
    assert float(eng.grads.abs().max()) == 0.0                          # zero_grad fused
    for i, n in enumerate(names):
        got, ref = summarize(eng.view(n) - before[n]), g["adam_summary"][i]
        assert abs(got[2] - ref[2]) <= 2e-2 * ref[2] + 1e-8 * np.sqrt(eng.view(n).numel()), (n, got[2], ref[2])
        assert np.abs(got[3:] - ref[3:]).max() <= 2e-2 * np.abs(ref[3:]).max() + 1e-8, n
    for n in orc.FT_NO_GRAD:                                           # BertAdam skips p.grad is None: not even weight decay
        assert torch.equal(eng.view(n), before[n]), n
    # optimizer checkpoint (finetune.py:484-486 / :396-402): schedule position + Adam moments survive a round trip
    sd = opt.state_dict()
    m_ref, v_ref = eng.adam_m.clone(), eng.adam_v.clone()
    assert float(v_ref.abs().max()) > 0 and sd["state"]["step"] == 3
    eng.adam_m.zero_(); eng.adam_v.zero_()
    opt2 = BertAdam([{"params": [p for p in model.parameters()], "weight_decay": 0.01}], lr=float(g["adam_lr"]),
                    warmup=float(g["adam_warmup"]), t_total=int(g["adam_t_total"]))
    opt2.load_state_dict(sd)
    assert opt2.state["step"] == 3 and opt2.scheduled_lr() == opt.scheduled_lr()
    assert torch.equal(eng.adam_m, m_ref) and torch.equal(eng.adam_v, v_ref)


@pytest.mark.parametrize("name", TINY)
def test_finetune_step_bf16(name):
    g, cfg, batch, params, model, out = run(name, "bf16", explicit_mask=False)    # compact (mode, t_len) fast path
    assert abs(out["loss"] - float(g["loss"])) <= 1e-2 * abs(float(g["loss"])), (out["loss"], float(g["loss"]))
    eng = model.engine()
    ref = orc.finetune_loss_and_grads(params, batch, cfg, feats=oracle_feats(params, batch))["grads"]
    scale = max(float(v.norm()) for v in ref.values())
    for n, r in ref.items():
        got = eng.view(n, eng.grads).float().cpu().flatten().double()
        r = r.flatten().double()
        if float(r.norm()) <= 1e-4 * scale:
            assert float(got.norm()) <= 1e-3 * scale, n
            continue
        cos = float(got @ r / (got.norm() * r.norm()))
        assert cos >= 0.95, "grad cosine %s: %.5f" % (n, cos)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_finetune_step_bert_base_L512(precision, tol):
    """BERT-base, 256 regions + 253 report tokens (L = 512), B = 2: loss vs the reference fixture; forward() == the loss the
    training step reports (dropout 0); fp32 gradient norms vs the fixture"""
    g, cfg, batch, params, model, out = run("finetune_base_s2s", precision, explicit_mask=(precision == "fp32"))
    assert abs(out["loss"] - float(g["loss"])) <= tol * abs(float(g["loss"])), (out["loss"], float(g["loss"]))
    eng = model.engine()
    if precision == "fp32":
        for i, n in enumerate(str(x) for x in g["grad_names"]):
            got, ref = summarize(eng.view(n, eng.grads)), g["grad_summary"][i]
            if not n.endswith("attention.self.key.bias"):
                assert abs(got[2] - ref[2]) <= 5e-3 * max(ref[2], 1e-6), "grad norm %s: got %.6e ref %.6e" % (n, got[2], ref[2])
    t = lambda k: torch.as_tensor(batch[k])
    loss, dummy = model(None, None, t("input_ids"), t("segment_ids"), None, t("masked_ids"), masked_pos=t("masked_pos"),
                        masked_weights=t("masked_weights"), drop_worst_ratio=0, mode=t("mode"), t_len=t("t_len"),
                        feats=oracle_feats(params, batch))
    assert abs(float(loss) - out["loss"]) <= 1e-5 * abs(out["loss"]) and float(dummy) == 0.0
    with pytest.raises(Exception):                                     # int(2 * 0.1) = 0 samples kept
        model(None, None, t("input_ids"), t("segment_ids"), None, t("masked_ids"), masked_pos=t("masked_pos"),
              masked_weights=t("masked_weights"), drop_worst_ratio=0.9, mode=t("mode"), t_len=t("t_len"),
              feats=oracle_feats(params, batch))


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_finetune_drop_worst(precision, tol):
    """Luo's drop-worst (model.py:1003-1010) at ratio 0.3, B = 5: int(3.5) = 3 samples with the smallest loss kept; loss and
    gradients vs the reference fixture / the oracle, and forward() reports the same loss"""
    g, cfg, batch, params, model, out = run("finetune_tiny_s2s_dropworst", precision, explicit_mask=False,
                                            drop_worst_ratio=float(load_golden("finetune_tiny_s2s_dropworst")[0]["drop_worst_ratio"]))
    ratio = float(g["drop_worst_ratio"])
    assert ratio == 0.3 and len(g["kept_samples"]) == 3
    assert abs(out["loss"] - float(g["loss"])) <= tol * abs(float(g["loss"])), (out["loss"], float(g["loss"]))
    eng = model.engine()
    feats = oracle_feats(params, batch)
    ref = orc.finetune_loss_and_grads(params, batch, cfg, feats=feats, drop_worst_ratio=ratio)["grads"]
    full = orc.finetune_loss_and_grads(params, batch, cfg, feats=feats)["grads"]
    scale = max(float(v.norm()) for v in ref.values())
    n_checked = 0
    for n, r in ref.items():
        got = eng.view(n, eng.grads).float().cpu().flatten().double()
        r = r.flatten().double()
        if float(r.norm()) <= 1e-4 * scale:
            continue
        n_checked += 1
        cos = float(got @ r / (got.norm() * r.norm()))
        assert cos >= (0.9999 if precision == "fp32" else 0.95), "grad cosine %s: %.5f" % (n, cos)
        if precision == "fp32":
            assert abs(float(got.norm()) - float(r.norm())) <= 3e-3 * float(r.norm()), n
    assert n_checked >= 30
    # the dropped samples matter: the all-samples gradient of the decoder bias differs visibly from the kept-only one
    n = "mlm.predictions.bias"
    assert float((ref[n] - full[n]).norm()) > 0.05 * float(full[n].norm())
    t = lambda k: torch.as_tensor(batch[k])
    loss, _ = model(None, None, t("input_ids"), t("segment_ids"), None, t("masked_ids"), masked_pos=t("masked_pos"),
                    masked_weights=t("masked_weights"), drop_worst_ratio=ratio, mode=t("mode"), t_len=t("t_len"), feats=feats)
    assert abs(float(loss) - out["loss"]) <= 1e-5 * abs(out["loss"])
    small = make_model(cfg, precision, params, max_batch=2)            # 5 samples in 3 micro-batches: cannot rank the batch
    with pytest.raises(Exception, match="one micro-batch"):
        small.finetune_step(None, t("input_ids"), t("segment_ids"), None, t("masked_ids"), t("masked_pos"), t("masked_weights"),
                            mode=t("mode"), t_len=t("t_len"), feats=feats, drop_worst_ratio=ratio)


@pytest.mark.parametrize("L,A", [(512, 258), (39, 18)])
def test_finetune_mask_modes_bit_exact_and_classified(L, A):
    """mv_attn_mask_dump for the fine-tune variants == the oracle's construction; mv_mask_classify recovers (mode, t_len)
    from the explicit [L, L] masks (and falls back to the pre-training modes when there is no padding)"""
    from medvill_b200 import _lib

    T = L - A
    tl = np.array([1, 2, T // 2, T - 1, T, T], dtype=np.int32)
    modes = np.array([orc.MODE_S2S_FT, orc.MODE_BAR_FT, orc.MODE_S2S_FT, orc.MODE_BAR_FT, orc.MODE_S2S_FT, orc.MODE_BAR_FT], dtype=np.uint8)
    want = np.stack([orc.attention_mask(int(m), A, L, int(t)) for m, t in zip(modes, tl)])
    B = len(tl)
    d_mode, d_tl = torch.from_numpy(modes).cuda(), torch.from_numpy(tl).cuda()
    out = torch.empty(B, L, L, dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib().mv_attn_mask_dump(_lib.ptr(d_mode), _lib.ptr(d_tl), B, A, L, _lib.ptr(out), _lib.stream_ptr()))
    assert np.array_equal(out.cpu().numpy().astype(np.int64), want)
    mask = torch.from_numpy(want).cuda()
    c_mode = torch.empty(B, dtype=torch.uint8, device="cuda")
    c_tl = torch.empty(B, dtype=torch.int32, device="cuda")
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().mv_mask_classify(_lib.ptr(mask), 3, B, A, L, _lib.ptr(c_mode), _lib.ptr(c_tl), _lib.ptr(bad), _lib.stream_ptr()))
    assert int(bad) == 0
    got_mode, got_tl = c_mode.cpu().numpy(), c_tl.cpu().numpy()
    for i in range(B):
        if tl[i] < T:
            assert got_mode[i] == modes[i] and got_tl[i] == tl[i], i
        else:       # no padding: identical to the pre-training Seq2Seq / BAR mask
            assert got_mode[i] == (orc.MODE_S2S if modes[i] == orc.MODE_S2S_FT else orc.MODE_BAR), i
