"""The drop-in surface under torch.autograd: the REFERENCE's own step body (models/train_origin.py:106-131, restated below because
/root/reference does not exist on the GPU box) and the retrieval model's composition (Downstream_task/Retrieval/retrieval.py:29-32)
run unmodified on the repo's CXRBERT, and give the same losses / gradients as the fused `pretrain_step` and the reference fixtures."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from tests.test_model_gpu import make_model
from tests.util import golden_batch, load_golden

pytestmark = pytest.mark.gpu


def _inputs(batch, dev="cuda:0"):
    t = lambda k: torch.as_tensor(batch[k]).to(dev)
    return dict(cls_tok=t("cls_tok"), input_ids=t("input_ids"), attn_masks=t("attn_masks"), segment=t("segment"), img=batch["image"].to(dev),
                sep_tok=t("sep_tok"), txt_labels=t("txt_labels"), is_aligned=t("is_aligned"))


def reference_step_body(model, optimizer, x):
    """models/train_origin.py:106-131, line for line (mlm_task and itm_task both on, the only case the CLI reaches)"""
    mlm_criterion = nn.CrossEntropyLoss(ignore_index=-100)          # :62
    itm_criterion = nn.CrossEntropyLoss()                           # :63
    mlm_output, itm_output = model(x["cls_tok"], x["input_ids"], x["attn_masks"], x["segment"], x["img"], x["sep_tok"])   # :106
    mlm_loss = mlm_criterion(mlm_output.transpose(1, 2), x["txt_labels"])                                               # :120
    itm_loss = itm_criterion(itm_output, x["is_aligned"])                                                               # :123
    loss = itm_loss + mlm_loss                                                                                          # :126
    optimizer.zero_grad()                                                                                               # :129
    loss.backward()                                                                                                     # :130
    return loss, mlm_loss, itm_loss, mlm_output, itm_output


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 2e-2)])
def test_reference_training_loop_runs_on_the_dropin_model(precision, tol):
    g, cfg = load_golden("tiny_mixed")
    batch = golden_batch(g, cfg)
    model, _ = make_model(cfg, precision)
    model.train()
    model.enc.img_encoder.region_idx_override = batch["region_idx"]
    x = _inputs(batch)
    # the fused step (forward + CE + backward, no optimizer step) on the same weights: the gradient to match
    t = lambda k: torch.as_tensor(batch[k])
    eng = model.engine(int(g["B"]))
    eng.zero_grads()
    model.pretrain_step(t("cls_tok"), t("input_ids"), t("txt_labels"), t("attn_masks"), batch["image"], t("segment"), t("is_aligned"),
                        t("sep_tok"), optimizer_step=False)
    torch.cuda.synchronize()
    fused = eng.grads.clone()
    eng.zero_grads()

    optimizer = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-3, eps=1e-6, weight_decay=0.0)
    loss, mlm_loss, itm_loss, mlm_out, itm_out = reference_step_body(model, optimizer, x)
    assert mlm_out.shape == (int(g["B"]), cfg.L, cfg.vocab) and mlm_out.grad_fn is not None and itm_out.grad_fn is not None
    ltol = 1e-4 if precision == "fp32" else 1e-2
    assert abs(float(loss) - float(g["loss"])) <= ltol * float(g["loss"])
    assert abs(float(mlm_loss) - float(g["mlm_loss"])) <= ltol * float(g["mlm_loss"])
    got = eng.grads
    err = float((got - fused).norm() / fused.norm())
    print("%s: autograd-path gradient vs fused pretrain_step: rel-L2 %.3e" % (precision, err))
    assert err <= tol
    # every parameter's .grad is the arena view the optimizer will read
    w = model.enc.encoder.layer[0].attention.self.query.weight
    assert w.grad is not None and w.grad.data_ptr() == eng.view("enc.encoder.layer.0.attention.self.query.weight", eng.grads).data_ptr()
    before = w.detach().clone()
    optimizer.step()                                                                                                    # :131
    assert float((w.detach() - before).abs().max()) > 0
    # second step: the forward sees the stepped weights (bf16 operand shadow refreshed), zero_grad(set_to_none) is handled
    loss2, *_ = reference_step_body(model, optimizer, x)
    optimizer.step()
    assert np.isfinite(float(loss2)) and float(loss2) < float(loss)          # lr 1e-3 on a 3-sample batch: the loss drops


def test_retrieval_composition_enc_then_itm():
    """CXRBertForRetrieval.forward (Downstream_task/Retrieval/retrieval.py:29-32): `_, cls, _ = self.enc(...)`; `self.itm(cls)`"""
    g, cfg = load_golden("tiny_bar")
    batch = golden_batch(g, cfg)
    model, _ = make_model(cfg, "fp32")
    model.enc.img_encoder.region_idx_override = batch["region_idx"]
    x = _inputs(batch)

    class CXRBertForRetrieval(nn.Module):                      # retrieval.py:12-32 with the repo's modules
        def __init__(self, inner):
            super().__init__()
            self.enc, self.itm = inner.enc, inner.itm

        def forward(self, cls_tok, input_txt, attn_mask, segment, input_img, sep_tok):
            _, cls, _ = self.enc(cls_tok, input_txt, attn_mask, segment, input_img, sep_tok)
            return self.itm(cls)

    r = CXRBertForRetrieval(model)
    model.train()                                             # train-mode BatchNorm, as the fixture's reference run
    with torch.no_grad():
        logits = r(x["cls_tok"], x["input_ids"], x["attn_masks"], x["segment"], x["img"], x["sep_tok"])
    assert np.abs(logits.cpu().numpy() - g["itm_logits"]).max() <= 2e-4
    # with gradients: ITM loss through enc -> itm == ITM loss through the joint forward
    eng = model.engine()
    crit = nn.CrossEntropyLoss()
    model.zero_grad(set_to_none=True)
    _, itm = model(x["cls_tok"], x["input_ids"], x["attn_masks"], x["segment"], x["img"], x["sep_tok"])
    crit(itm, x["is_aligned"]).backward()
    torch.cuda.synchronize()
    joint = eng.grads.clone()
    assert float(joint.norm()) > 0
    model.zero_grad(set_to_none=True)
    crit(r(x["cls_tok"], x["input_ids"], x["attn_masks"], x["segment"], x["img"], x["sep_tok"]), x["is_aligned"]).backward()
    torch.cuda.synchronize()
    err = float((eng.grads - joint).norm() / joint.norm())
    print("enc -> itm composition vs joint forward, ITM gradient: rel-L2 %.3e" % err)
    assert err <= 1e-4
    assert model.itm.linear.weight.grad.data_ptr() == eng.view("itm.linear.weight", eng.grads).data_ptr()
