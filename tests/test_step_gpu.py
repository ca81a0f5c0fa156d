"""GPU parity of the whole pre-training step (forward, losses, metrics, backward, AdamW) through the C ABI against
the golden fixtures produced by the real reference (tests/golden, oracle/make_golden.py) and the CPU oracle.

Tolerances (BASELINE.json north_star): fp32 check mode 1e-4 relative on losses/logits; bf16 1e-2 relative.
MLM token selection (labelled rows, argmax) and masks are bit-exact.
"""
import numpy as np
import pytest
import torch

from oracle import medvill_oracle as orc
from tests.util import dims_from_cfg, golden_batch, load_golden, oracle_feats, summarize

pytestmark = pytest.mark.gpu

TINY = ["tiny_bar", "tiny_s2s", "tiny_noncross", "tiny_bidir", "tiny_mixed"]


def run_step(name, precision):
    import medvill_b200 as m

    g, cfg = load_golden(name)
    batch = golden_batch(g, cfg)
    params = orc.synth_params(cfg, seed=0)
    feats = oracle_feats(params, batch)
    eng = m.PretrainEngine(dims_from_cfg(cfg), "cuda:0", precision=precision, max_batch=int(g["B"]))
    eng.load_params(params)
    b = eng.make_batch(cls_tok=batch["cls_tok"], input_ids=batch["input_ids"], segment=batch["segment"], sep_tok=batch["sep_tok"],
                       mode=batch["mode"], t_len=batch["t_len"], region_idx=batch["region_idx"], feats=feats,
                       txt_labels=batch["txt_labels"], is_aligned=batch["is_aligned"], seed=1, train=True)
    assert np.array_equal(b.lab_rows.cpu().numpy(), g["lab_rows"][:, 0] * cfg.L + g["lab_rows"][:, 1])   # bit-exact token selection
    eng.zero_grads()
    eng.stats_reset()
    eng.forward(b)
    st = eng.read_stats()
    out = dict(g=g, cfg=cfg, eng=eng, b=b, st=st, params=params)
    out["mlm_loss"] = st["mlm_loss_sum"] / b.n_lab
    out["itm_loss"] = st["itm_loss_sum"] / b.B
    out["itm_logits"] = eng.itm_logits(b.B).numpy()
    out["row_lse"] = eng.peek("row_lse", shape=(b.n_lab,), dtype=torch.float32).numpy()
    out["row_argmax"] = eng.peek("row_argmax", shape=(b.n_lab,), dtype=torch.int32).numpy()
    ld = eng.layout["vocab_padded"]
    logits = eng.peek("logits", shape=(b.n_lab, ld), dtype=torch.float32).numpy()
    out["lab_logits"] = logits[:, g["lab_cols"]]
    eng.backward(b)
    torch.cuda.synchronize()
    return out


def check_forward(o, tol, tol_max=None):
    """losses: |d| / |ref| <= tol.  logits: relative Frobenius error <= tol and worst element <= tol_max * max|ref|
    (tol_max defaults to tol; bf16 runs allow 3x for the single worst of ~5e3 probed logits after 12 bf16 layers)."""
    g = o["g"]
    tol_max = tol if tol_max is None else tol_max
    assert abs(o["mlm_loss"] - float(g["mlm_loss"])) <= tol * abs(float(g["mlm_loss"])), (o["mlm_loss"], float(g["mlm_loss"]))
    assert abs(o["itm_loss"] - float(g["itm_loss"])) <= tol * max(1.0, abs(float(g["itm_loss"]))), (o["itm_loss"], float(g["itm_loss"]))
    diff = o["lab_logits"].astype(np.float64) - g["lab_logits"].astype(np.float64)
    rel_l2 = np.linalg.norm(diff) / np.linalg.norm(g["lab_logits"].astype(np.float64))
    rel_max = np.abs(diff).max() / np.abs(g["lab_logits"]).max()
    print("logits rel-L2 %.3e  worst/max %.3e  mlm_loss %.6f (ref %.6f) itm_loss %.6f (ref %.6f)" % (
        rel_l2, rel_max, o["mlm_loss"], float(g["mlm_loss"]), o["itm_loss"], float(g["itm_loss"])))
    assert rel_l2 <= tol, rel_l2
    assert rel_max <= tol_max, rel_max
    assert np.abs(o["itm_logits"] - g["itm_logits"]).max() <= tol * max(1.0, np.abs(g["itm_logits"]).max())
    assert np.abs(o["row_lse"] - g["lab_lse"]).max() <= tol * np.abs(g["lab_lse"]).max()
    assert o["st"]["itm_correct"] == int(g["itm_correct"])


def check_grads(o, tol_norm, tol_probe):
    g, eng = o["g"], o["eng"]
    names = [str(n) for n in g["grad_names"]]
    worst = 0.0
    scale = max(r[2] for r in g["grad_summary"])
    for i, n in enumerate(names):
        got = summarize(eng.view(n, eng.grads))
        ref = g["grad_summary"][i]
        if n.endswith("attention.self.key.bias"):
            # analytically zero (softmax is invariant to a per-query constant): both sides hold rounding noise only
            assert got[2] <= 1e-4 * scale and ref[2] <= 1e-4 * scale, (n, got[2], ref[2])
            continue
        nrm = max(ref[2], 1e-6)       # analytically-zero gradients (key biases) only hold rounding noise
        e_norm = abs(got[2] - ref[2]) / nrm
        e_probe = np.abs(got[3:] - ref[3:]).max() / max(np.abs(ref[3:]).max(), 1e-3 * nrm, 1e-8)
        worst = max(worst, e_norm)
        assert e_norm <= tol_norm, "grad norm %s: got %.6e ref %.6e" % (n, got[2], ref[2])
        assert e_probe <= tol_probe, "grad probes %s: got %s ref %s" % (n, got[3:], ref[3:])
    return worst


@pytest.mark.parametrize("name", TINY)
def test_step_fp32_check_mode_tiny(name):
    o = run_step(name, "fp32")
    check_forward(o, 1e-4)
    assert np.array_equal(o["row_argmax"], o["g"]["lab_argmax"])
    assert o["st"]["mlm_correct"] == int(o["g"]["mlm_correct"])
    check_grads(o, 2e-3, 5e-3)
    # one AdamW step (HF-3.x formula, lr 1e-5): parameter deltas
    eng, g = o["eng"], o["g"]
    before = {n: eng.view(n).clone() for n in eng.pmap}
    eng.adamw_step(lr=1e-5)
    torch.cuda.synchronize()
    for i, n in enumerate(str(x) for x in g["grad_names"]):
        got = summarize(eng.view(n) - before[n])
        ref = g["adamw_summary"][i]
        # key-bias gradients are analytically zero, so Adam only normalises rounding noise there: absolute floor
        assert abs(got[2] - ref[2]) <= 2e-2 * ref[2] + 1e-8 * np.sqrt(eng.view(n).numel()), (n, got[2], ref[2])
    assert float(eng.grads.abs().max()) == 0.0     # optimizer.zero_grad() fused into the step


@pytest.mark.parametrize("name", TINY)
def test_step_bf16_tiny(name):
    o = run_step(name, "bf16")
    check_forward(o, 1e-2, 3e-2)
    # measured (r02, fp32 residual stream): tensors carrying >= 1 % of the largest gradient norm: cosine >= 0.99995, norm within
    # 0.3 %; the small ITM-head / pooler tensors of the B = 4 mixed batch (near-tied ITM logits): 0.972 / 6 %
    w = check_grads_vs_oracle(o, 0.9995, 0.01, cos_min_small=0.96, norm_tol_small=0.08)
    print("worst grad cosine %.5f" % w)


def test_step_fp32_config1():
    o = run_step("config1_bar", "fp32")
    check_forward(o, 1e-4)
    assert np.array_equal(o["row_argmax"], o["g"]["lab_argmax"])
    check_grads(o, 3e-3, 1e-2)


def check_grads_vs_oracle(o, cos_min, norm_tol, cos_min_small=None, norm_tol_small=None):
    """bf16 gradients, tensor by tensor, against the CPU oracle's full gradients (the oracle itself is pinned to the
    reference to 1e-5): cosine similarity and norm ratio.  Element probes are meaningless in bf16 for entries far below
    the tensor's scale (LayerNorm-backward cancellation), so this is the bf16 gradient gate.  Tensors whose gradient norm is
    below 1 % of the largest tensor's (biases / LayerNorm gains whose terms nearly cancel) get `cos_min_small`."""
    eng, cfg, params = o["eng"], o["cfg"], o["params"]
    batch = golden_batch(o["g"], cfg)
    ref = orc.loss_and_grads(params, batch, cfg, feats=oracle_feats(params, batch))["grads"]
    scale = max(float(v.norm()) for v in ref.values())
    cos_min_small = cos_min if cos_min_small is None else cos_min_small
    norm_tol_small = norm_tol if norm_tol_small is None else norm_tol_small
    rows = []
    for n, r in ref.items():
        got = eng.view(n, eng.grads).float().cpu().flatten().double()
        r = r.flatten().double()
        if float(r.norm()) <= 1e-4 * scale:            # analytically-zero gradients (key biases)
            assert float(got.norm()) <= 1e-3 * scale, n
            continue
        cos = float(got @ r / (got.norm() * r.norm()))
        rows.append((cos, abs(float(got.norm()) / float(r.norm()) - 1.0), float(r.norm()) / scale, n))
    rows.sort()
    for cos, dn, rel, n in rows[:3]:
        print("  low-cosine tensor %-60s cos %.5f  norm dev %.4f  |g| / max|g| %.2e" % (n, cos, dn, rel))
    print("  worst norm deviation %.4f" % max(r[1] for r in rows))
    for cos, dn, rel, n in rows:
        assert cos >= (cos_min if rel >= 1e-2 else cos_min_small), "grad cosine %s: %.5f" % (n, cos)
        assert dn <= (norm_tol if rel >= 1e-2 else norm_tol_small), "grad norm %s: deviation %.4f" % (n, dn)
    return rows[0][0]


def autocast_logit_error(o):
    """Calibration: the SAME oracle arithmetic run by PyTorch on the GPU under bf16 autocast, measured against the
    fp32 reference fixture.  It is the error floor any honest bf16 tensor-core implementation of 12 BERT layers has
    (operand rounding of ~72 chained contractions), so our bf16 path is required to stay within 1.5x of it."""
    g, cfg, params = o["g"], o["cfg"], o["params"]
    batch = golden_batch(g, cfg)
    dev = torch.device("cuda:0")
    p = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in params.items()}
    b = {k: (torch.as_tensor(v).to(dev) if k != "mode" else v) for k, v in batch.items()}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        feats = oracle_feats(params, batch).to(dev)
        logits, _ = orc.forward(p, b, cfg, feats=feats)
    rows = g["lab_rows"]
    lab = logits.float()[rows[:, 0], rows[:, 1]][:, torch.as_tensor(g["lab_cols"]).to(dev)].cpu().numpy().astype(np.float64)
    ref = g["lab_logits"].astype(np.float64)
    return np.linalg.norm(lab - ref) / np.linalg.norm(ref)


def test_step_bf16_config1():
    """BERT-base, L=436, B=2 (BASELINE.json configs[0] shapes) in the production bf16 mode.  Losses and logits: 1e-2
    relative (north_star); logits also <= 1.3x the PyTorch-autocast bf16 floor (see autocast_logit_error) — met because
    the residual stream (pre-LayerNorm sums and the LayerNorm outputs that feed the next residual add) is kept in fp32."""
    o = run_step("config1_bar", "bf16")
    check_forward(o, 1e-2, 3e-2)
    g = o["g"]
    assert abs(o["mlm_loss"] - float(g["mlm_loss"])) <= 1e-2 * float(g["mlm_loss"])
    assert abs(o["mlm_loss"] + o["itm_loss"] - float(g["loss"])) <= 1e-2 * float(g["loss"])
    diff = o["lab_logits"].astype(np.float64) - g["lab_logits"].astype(np.float64)
    ours = np.linalg.norm(diff) / np.linalg.norm(g["lab_logits"].astype(np.float64))
    floor = autocast_logit_error(o)
    print("bf16 logits rel-L2: ours %.3e, torch autocast floor %.3e" % (ours, floor))
    assert ours <= 1.3 * floor        # same rounding points as autocast: bf16 GEMM operands, fp32 residual stream
    w = check_grads_vs_oracle(o, 0.999, 0.02)        # measured: worst cosine 0.9996, worst norm deviation 1.0 %
    print("worst grad cosine %.5f" % w)
