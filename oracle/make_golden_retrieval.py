"""TEST INFRASTRUCTURE — pins the retrieval-scoring restatement (oracle/medvill_oracle.py: retrieval_pair,
retrieval_scores, retrieval_rank_metrics) against the real reference and writes tests/golden/retrieval_tiny.npz.

Run in the build container (needs /root/reference):   python oracle/make_golden_retrieval.py
  * imports the UNMODIFIED Downstream_task/Retrieval/retrieval.py (CXRBertForRetrieval) under oracle/ref_shim.py, loads the
    oracle's deterministic weights (orc.retrieval_params: non-trivial BatchNorm running statistics, scaled ITM head), runs
    model.eval() forward + nn.Softmax(dim=1)(logits)[:, 1] on every (image, report) pair of a small grid and asserts
    agreement with the restatement;
  * executes the reference's own compute_ranks / compute_recall_precision / compute_mrr / evaluate (function bodies
    taken from Downstream_task/Retrieval/full_dset_retrieval.py:250-339 at generation time — that module cannot be imported
    whole: it parses argv and pulls wandb / fuzzywuzzy at import) on seeded similarity tables including ties, saturated
    scores and groups without an aligned candidate, and asserts the restated metrics are identical.
"""
import ast
import contextlib
import io
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import medvill_oracle as orc  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.make_golden import _RandpermInject, ref_args  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def build_reference_retrieval(cfg, params):
    from transformers import BertConfig

    ref_shim.load_reference_models()           # registers models.cxrbert_origin for retrieval.py's import
    kw = dict(vocab_size=cfg.vocab, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
              intermediate_size=cfg.inter, max_position_embeddings=cfg.max_pos, type_vocab_size=cfg.type_vocab,
              hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, layer_norm_eps=cfg.ln_eps, hidden_act="gelu")
    ref_shim.set_bert_config(**kw)
    mod = ref_shim._load("ref_retrieval", "Downstream_task/Retrieval/retrieval.py")
    config = BertConfig(attn_implementation="eager", **kw)
    model = mod.CXRBertForRetrieval(config, ref_args(cfg, weight_load=False))
    sd = model.state_dict()
    new = {}
    for k in sd:
        ck = orc.canonical_key(k)
        new[k] = params[ck].clone() if ck in params else sd[k]
        assert ck in params or k.endswith("position_ids") or k.endswith("token_type_ids"), k
    model.load_state_dict(new, strict=True)
    model.eval()                               # full_dset_retrieval.py:462
    return model


def reference_metric_functions():
    src = open(os.path.join(ref_shim.REF_ROOT, "Downstream_task/Retrieval/full_dset_retrieval.py")).read()
    want = {"compute_ranks", "compute_recall_precision", "compute_mrr", "evaluate"}
    tree = ast.parse(src)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    assert {n.name for n in body} == want
    ns = {"np": np}
    exec(compile(ast.Module(body=body, type_ignores=[]), "full_dset_retrieval.py", "exec"), ns)
    return ns


def metric_cases():
    rng = np.random.RandomState(2024)
    cases = []
    for group, n_groups, kind in ((12, 7, "plain"), (15, 5, "ties"), (10, 6, "saturated"), (11, 4, "none_aligned")):
        sims = rng.rand(n_groups, group).astype(np.float32)
        labels = (rng.rand(n_groups, group) < 0.2).astype(np.int64)
        labels[np.arange(n_groups), rng.randint(0, group, n_groups)] = 1
        if kind == "ties":
            sims = np.round(sims * 4) / 4
        if kind == "saturated":
            sims[sims > 0.6] = 1.0
            sims[sims < 0.3] = 0.0
        if kind == "none_aligned":
            labels[1] = 0
        idx = rng.permutation(n_groups * group).reshape(n_groups, group)
        cases.append((kind, group, sims, labels, idx))
    return cases


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.manual_seed(0)
    cfg = orc.Cfg(**orc.TINY)
    params = orc.retrieval_params(cfg, seed=0)
    model = build_reference_retrieval(cfg, params)
    n_img, n_txt, seed = 3, 5, 31
    nrng = np.random.RandomState(seed)
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(n_img, 3, cfg.img_size, cfg.img_size, generator=g)
    lens = [1, cfg.seq_len, 7, 13, cfg.seq_len + 9]            # shortest, exactly full, ragged, truncated
    reports = [nrng.randint(200, cfg.vocab, size=t).tolist() for t in lens]
    region_idx = np.sort(nrng.permutation(cfg.grid)[:cfg.num_image_embeds]).astype(np.int64)
    pairs = [orc.retrieval_pair(r, cfg) for r in reports]
    stack = lambda k: np.stack([p[k] for p in pairs])
    scores_ref = np.zeros((n_img, n_txt), dtype=np.float32)
    logits_ref = np.zeros((n_img, n_txt, 2), dtype=np.float32)
    scores_orc = np.zeros_like(scores_ref)
    with torch.no_grad():
        fmap = orc.resnet50_trunk(params, images, bn_train=False)
        feats = torch.flatten(fmap, start_dim=2).transpose(1, 2).contiguous()
        for i in range(n_img):
            t = lambda k: torch.as_tensor(stack(k))
            img = images[i:i + 1].expand(n_txt, -1, -1, -1)
            with _RandpermInject(region_idx, cfg.grid):
                lg = model(t("cls_tok"), t("input_ids"), t("attn_masks"), t("segment"), img, t("sep_tok"))
            logits_ref[i] = lg.numpy()
            scores_ref[i] = torch.nn.Softmax(dim=1)(lg)[:, 1].numpy()
            batch = dict(cls_tok=stack("cls_tok"), input_ids=stack("input_ids"), attn_masks=stack("attn_masks"),
                         segment=stack("segment"), sep_tok=stack("sep_tok"), region_idx=region_idx)
            scores_orc[i] = orc.retrieval_scores(params, batch, cfg, feats=feats[i:i + 1].expand(n_txt, -1, -1)).numpy()
    err = float(np.abs(scores_ref - scores_orc).max())
    print("[pin] retrieval scores: max |ref - oracle| = %.2e over %d pairs (score range %.3f..%.3f)" % (
        err, n_img * n_txt, scores_ref.min(), scores_ref.max()))
    assert err < 2e-6, "retrieval score mismatch"

    ns = reference_metric_functions()
    out = dict(cfg=json.dumps(cfg.__dict__), seed=seed, n_img=n_img, n_txt=n_txt, lens=np.asarray(lens), region_idx=region_idx,
               input_ids=stack("input_ids"), attn_masks=stack("attn_masks"), t_len=np.asarray([p["t_len"] for p in pairs]),
               scores=scores_ref, logits=logits_ref)
    for ci, (kind, group, sims, labels, idx) in enumerate(metric_cases()):
        args = types.SimpleNamespace(eval_len_size=group, i2t=True, t2i=False)
        results = [torch.tensor(float(s)) for s in sims.reshape(-1)]
        with contextlib.redirect_stdout(io.StringIO()):
            ev, aligned, mrr, rp = ns["evaluate"](args, results, labels.reshape(-1).tolist(), idx.reshape(-1).tolist())
            ranks, _, _ = ns["compute_ranks"](args, results, labels.reshape(-1).tolist(), idx.reshape(-1).tolist())
        mine = orc.retrieval_rank_metrics(sims, labels, idx, group)
        assert mine["ranks"] == [int(r) for r in ranks], kind
        assert [[int(a), int(b)] for a, b in mine["aligned"]] == [[int(a), int(b)] for a, b in aligned], kind
        assert mine["hits"] == ev["i2t_retrieval"], kind
        assert np.isclose(mine["mrr"], mrr, rtol=0, atol=0), kind
        for k in ("R@1", "R@5", "R@10"):
            a, b = mine["recall"][k], rp["i2t_recall"][k]
            assert a == b or (np.isnan(a) and np.isnan(b)), (kind, k, a, b)
            assert mine["precision"][k] == rp["i2t_precision"][k], kind
        print("[pin] rank metrics '%s': ranks %s mrr %.4f — identical" % (kind, mine["ranks"], mrr))
        p = "m%d_" % ci
        out.update({p + "kind": kind, p + "group": group, p + "sims": sims, p + "labels": labels, p + "idx": idx,
                    p + "ranks": np.asarray(ranks), p + "aligned": np.asarray(aligned), p + "mrr": mrr,
                    p + "hits": np.asarray([ev["i2t_retrieval"][k] for k in ("R@1", "R@5", "R@10")]),
                    p + "recall": np.asarray([rp["i2t_recall"][k] for k in ("R@1", "R@5", "R@10")], dtype=np.float64),
                    p + "precision": np.asarray([rp["i2t_precision"][k] for k in ("R@1", "R@5", "R@10")], dtype=np.float64)})
    out["n_metric_cases"] = len(metric_cases())
    np.savez_compressed(os.path.join(GOLDEN, "retrieval_tiny.npz"), **out)
    print("   wrote retrieval_tiny.npz")


if __name__ == "__main__":
    main()
