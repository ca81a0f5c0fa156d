"""CPU checks of the retrieval-scoring path (BASELINE.json configs[4]): the oracle restatement against the
fixture generated from the real reference (oracle/make_golden_retrieval.py), and the product's host-side logic (report
encoding, rank metrics) against the same fixture — bit-exact for the integer outputs."""
import types

import numpy as np
import torch

from oracle import medvill_oracle as orc
from tests.util import load_golden


def _pairs(g, cfg):
    nrng = np.random.RandomState(int(g["seed"]))
    gen = torch.Generator().manual_seed(int(g["seed"]))
    images = torch.randn(int(g["n_img"]), 3, cfg.img_size, cfg.img_size, generator=gen)
    reports = [nrng.randint(200, cfg.vocab, size=int(t)).tolist() for t in g["lens"]]
    region_idx = np.sort(nrng.permutation(cfg.grid)[:cfg.num_image_embeds]).astype(np.int64)
    assert np.array_equal(region_idx, g["region_idx"])
    return images, reports, region_idx


def test_oracle_reproduces_reference_retrieval_scores():
    g, cfg = load_golden("retrieval_tiny")
    images, reports, region_idx = _pairs(g, cfg)
    pairs = [orc.retrieval_pair(r, cfg) for r in reports]
    stack = lambda k: np.stack([p[k] for p in pairs])
    assert np.array_equal(stack("input_ids"), g["input_ids"]) and np.array_equal(stack("attn_masks"), g["attn_masks"])
    params = orc.retrieval_params(cfg, seed=0)
    with torch.no_grad():
        fmap = orc.resnet50_trunk(params, images, bn_train=False)
        feats = torch.flatten(fmap, start_dim=2).transpose(1, 2).contiguous()
        for i in range(int(g["n_img"])):
            batch = dict(cls_tok=stack("cls_tok"), input_ids=stack("input_ids"), attn_masks=stack("attn_masks"),
                         segment=stack("segment"), sep_tok=stack("sep_tok"), region_idx=region_idx)
            f = feats[i:i + 1].expand(len(pairs), -1, -1)
            lg = orc.retrieval_logits(params, batch, cfg, feats=f).numpy()
            sc = orc.retrieval_scores(params, batch, cfg, feats=f).numpy()
            assert np.abs(lg - g["logits"][i]).max() < 1e-5 * np.abs(g["logits"][i]).max()
            assert np.abs(sc - g["scores"][i]).max() < 2e-6


def test_product_report_encoding_matches_reference_layout():
    """data_processing (full_dset_retrieval.py:199-218): ids + [SEP], zero padding, 1-D mask, truncation to seq_len"""
    from medvill_b200.retrieval import data_processing

    g, cfg = load_golden("retrieval_tiny")
    _, reports, _ = _pairs(g, cfg)
    for j, r in enumerate(reports):
        p = data_processing(r, cfg.seq_len, cfg.num_image_embeds)
        assert np.array_equal(p["input_ids"].numpy(), g["input_ids"][j])
        assert np.array_equal(p["attn_masks"].numpy(), g["attn_masks"][j])
        assert p["t_len"] == int(g["t_len"][j]) and p["segment"].shape[0] == cfg.seq_len + 1
        # the 1-D mask is exactly the Bidirectional predicate with this t_len
        assert np.array_equal(p["attn_masks"].numpy(), orc.attention_mask(orc.MODE_BIDIR, cfg.A, cfg.L, p["t_len"])[0])


def test_rank_metrics_match_reference_functions():
    """compute_ranks / compute_recall_precision / compute_mrr / evaluate vs outputs of the reference's own functions:
    ties, saturated similarities and a group with no aligned candidate included; oracle and product both checked"""
    from medvill_b200.retrieval import compute_ranks, evaluate

    g, _ = load_golden("retrieval_tiny")
    for ci in range(int(g["n_metric_cases"])):
        p = "m%d_" % ci
        group, sims, labels, idx = int(g[p + "group"]), g[p + "sims"], g[p + "labels"], g[p + "idx"]
        mine = orc.retrieval_rank_metrics(sims, labels, idx, group)
        assert mine["ranks"] == g[p + "ranks"].tolist()
        args = types.SimpleNamespace(eval_len_size=group, i2t=True, t2i=False)
        results = [torch.tensor(float(s)) for s in sims.reshape(-1)]
        lab, ids = labels.reshape(-1).tolist(), idx.reshape(-1).tolist()
        ranks, t2i, aligned = compute_ranks(args, results, lab, ids)
        assert ranks == g[p + "ranks"].tolist() and t2i == []
        assert np.array_equal(np.asarray(aligned), g[p + "aligned"])
        ev, aligned2, mrr, rp = evaluate(args, results, lab, ids)
        assert [ev["i2t_retrieval"][k] for k in ("R@1", "R@5", "R@10")] == g[p + "hits"].tolist()
        assert mrr == float(g[p + "mrr"])
        for name, key in (("recall", "i2t_recall"), ("precision", "i2t_precision")):
            got = np.asarray([rp[key][k] for k in ("R@1", "R@5", "R@10")], dtype=np.float64)
            assert np.array_equal(got, g[p + name], equal_nan=True), (ci, name)
        # text-to-image direction only relabels the outputs
        args_t = types.SimpleNamespace(eval_len_size=group, i2t=False, t2i=True)
        i2t, ranks_t, _ = compute_ranks(args_t, results, lab, ids)
        assert i2t == [] and ranks_t == ranks


def test_retrieval_model_state_dict_has_the_reference_key_set():
    """CXRBertForRetrieval exposes enc.* and itm.* only (Downstream_task/Retrieval/retrieval.py:26-27)"""
    import medvill_b200  # noqa: F401
    from medvill_b200.config import BertConfig
    from medvill_b200.retrieval import CXRBertForRetrieval

    cfg = orc.Cfg(**orc.TINY)
    bc = BertConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
                    intermediate_size=cfg.inter, max_position_embeddings=cfg.max_pos, type_vocab_size=cfg.type_vocab)
    args = types.SimpleNamespace(img_hidden_sz=cfg.img_hidden, embedding_size=cfg.hidden, hidden_size=cfg.hidden, dropout_prob=0.0,
                                 img_encoder="random-pixel", num_image_embeds=cfg.num_image_embeds, img_size=cfg.img_size,
                                 seq_len=cfg.seq_len, weight_load=False)
    m = CXRBertForRetrieval(bc, args)
    keys = set(m.state_dict().keys())
    assert keys and all(k.startswith("enc.") or k.startswith("itm.") for k in keys)
    assert "itm.linear.weight" in keys and "enc.pooler.dense.weight" in keys and not any(k.startswith("mlm.") for k in keys)
    params = orc.retrieval_params(cfg, seed=0)
    m.load_state_dict({k: params[orc.canonical_key(k)] for k in keys}, strict=True)
    assert torch.equal(m.itm.linear.weight.detach(), params["itm.linear.weight"])
