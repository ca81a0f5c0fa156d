"""CPU: the oracle against the golden vectors written from the REAL reference (oracle/make_golden.py), closed-form
masks against the step-by-step construction, and the product's host-side dataset logic against both (bit-exact)."""
import json
import os
import random
import types

import numpy as np
import pytest
import torch

from oracle import medvill_oracle as orc
from tests.util import GOLDEN, golden_batch, load_golden


@pytest.mark.parametrize("name", ["tiny_bar", "tiny_s2s", "tiny_noncross", "tiny_bidir", "tiny_mixed"])
def test_oracle_reproduces_reference_fixture(name):
    g, cfg = load_golden(name)
    batch = golden_batch(g, cfg)
    params = orc.synth_params(cfg, seed=0)
    out = orc.loss_and_grads(params, batch, cfg)
    assert abs(out["loss"] - float(g["loss"])) < 2e-5 * float(g["loss"])
    assert abs(out["mlm_loss"] - float(g["mlm_loss"])) < 2e-5 * float(g["mlm_loss"])
    rows = g["lab_rows"]
    lab = out["logits"][rows[:, 0], rows[:, 1]][:, g["lab_cols"]].numpy()
    assert np.abs(lab - g["lab_logits"]).max() < 1e-5 * np.abs(g["lab_logits"]).max() + 1e-6
    assert np.abs(out["itm_logits"].numpy() - g["itm_logits"]).max() < 1e-5
    names = [str(n) for n in g["grad_names"]]
    for i, n in enumerate(names):
        gr = out["grads"][n].double()
        assert abs(float(gr.norm()) - g["grad_summary"][i][2]) <= 2e-3 * g["grad_summary"][i][2] + 1e-7, n
    itm_c, mlm_c, n_lab = orc.step_metrics(out["logits"], out["itm_logits"], batch)
    assert (itm_c, mlm_c, n_lab) == (int(g["itm_correct"]), int(g["mlm_correct"]), int(g["n_labelled"]))


@pytest.mark.parametrize("N,S", [(180, 253), (256, 253), (36, 128), (9, 20)])
def test_closed_form_masks_equal_reference_construction(N, S):
    A, T, L = N + 2, S + 1, N + S + 3
    for mode in (orc.MODE_BIDIR, orc.MODE_S2S, orc.MODE_BAR, orc.MODE_NONCROSS):
        for t_len in (2, T // 2, T):
            a = orc.attention_mask(mode, A, L, t_len)
            b = orc.dataset_mask_construction(mode, N, S, T, t_len)
            assert np.array_equal(a, b), (mode, t_len)
    # densities quoted in SURVEY.md §5.7 (L = 436)
    if (N, S) == (180, 253):
        assert abs(orc.attention_mask(orc.MODE_BAR, A, L, T).mean() - 0.8310) < 1e-3
        assert abs(orc.attention_mask(orc.MODE_S2S, A, L, T).mean() - 0.5878) < 1e-3
        assert abs(orc.attention_mask(orc.MODE_NONCROSS, A, L, T).mean() - 0.5136) < 1e-3


def test_random_word_known_answer():
    rng = random.Random(42)
    toks, labels = orc.random_word(list(range(1000, 1040)), rng, 30522)
    assert len(toks) == len(labels) == 40
    changed = [i for i, l in enumerate(labels) if l != -100]
    assert all(labels[i] == 1000 + i for i in changed)                  # label = original id
    assert all(toks[i] == 1000 + i for i in range(40) if i not in changed)
    # forced mask when nothing was selected (dataset_origin.py:204-207)
    class Never:
        def random(self):
            return 0.99
    t2, l2 = orc.random_word([5, 6, 7], Never(), 100)
    assert t2[0] == orc.MASK and l2 == [5, -100, -100]


VARIANTS = {
    "bar": dict(BAR_attn=True), "bidir": dict(BAR_attn=False), "bidir1d": dict(BAR_attn=False, attn_1d=True),
    "s2s": dict(Mixed=True, s2s_prob=1.0, bi_prob=0.0), "mixed": dict(Mixed=True, s2s_prob=0.75, bi_prob=0.25),
    "noncross": dict(BAR_attn=False, disturbing_mask=True),
}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_product_dataset_bit_exact_with_reference_fixture(variant, tmp_path):
    """medvill_b200.data.CXRDataset under random.seed(1234) == the real reference CXRDataset (fixture)."""
    from PIL import Image

    import medvill_b200  # noqa: F401
    from medvill_b200.data.dataset_origin import CXRDataset

    fx = np.load(os.path.join(GOLDEN, "dataset_seed1234.npz"))
    recs = [json.loads(str(r)) for r in fx["records"]]
    Image.fromarray(np.zeros((8, 8, 3), dtype=np.uint8)).save(tmp_path / "x.png")
    path = tmp_path / "train.jsonl"
    with open(path, "w") as f:
        for r in recs:
            f.write(json.dumps(r) + "\n")
    over = VARIANTS[variant]
    args = types.SimpleNamespace(bert_model="bert-base-scratch", num_image_embeds=180, seq_len=253, max_seq_len=512, img_channel=3,
                                 Mixed=False, BAR_attn=True, attn_1d=False, s2s_prob=1.0, bi_prob=0.0, disturbing_mask=False)
    for k, v in over.items():
        setattr(args, k, v)
    vocab = {str(i): i for i in range(30522)}
    vocab.update({"[PAD]": 0, "[UNK]": 100, "[CLS]": 101, "[SEP]": 102, "[MASK]": 103})
    ds = CXRDataset(str(path), lambda s: s.split(), lambda im: torch.zeros(3, 4, 4), args, vocab=vocab)
    random.seed(1234)
    A, L = 182, 436
    for i in range(len(ds)):
        cls_tok, ids, labels, attn, _img, seg, aligned, sep_tok, _prob = ds[i]
        assert np.array_equal(ids.numpy(), fx[variant + "_input_ids"][i]), i
        assert np.array_equal(labels.numpy(), fx[variant + "_txt_labels"][i]), i
        assert int(aligned) == int(fx[variant + "_is_aligned"][i])
        mode, t_len = int(fx[variant + "_mode"][i]), int(fx[variant + "_t_len"][i])
        if attn.dim() == 2:
            assert np.array_equal(attn.numpy(), orc.attention_mask(mode, A, L, t_len)), (variant, i)
        else:
            assert np.array_equal(attn.numpy(), np.asarray([1] * (A + t_len) + [0] * (L - A - t_len)))
        assert int(cls_tok) == 101 and int(sep_tok) == 102 and seg.tolist() == [1] * 254
    # compact form carries exactly (mode, t_len)
    args.compact_masks = True
    random.seed(1234)
    for i in range(4):
        attn = ds[i][3]
        assert attn.tolist() == [int(fx[variant + "_mode"][i]), int(fx[variant + "_t_len"][i])]


def test_adamw_restatement_matches_closed_form():
    p = {"w": torch.tensor([1.0, -2.0, 3.0])}
    g = {"w": torch.tensor([0.5, -0.25, 0.0])}
    out = orc.adamw_step(dict(p), g, {}, lr=1e-2, step=1)
    m, v = 0.1 * g["w"], 0.001 * g["w"] ** 2
    step = 1e-2 * (1 - 0.999) ** 0.5 / (1 - 0.9)
    assert torch.allclose(out["w"], p["w"] - step * m / (v.sqrt() + 1e-6), atol=1e-7)
