"""2-GPU data-parallel equivalence (needs >= 2 GPUs; `gpurun --gpus 2 -- python -m pytest tests/test_ddp_gpu.py -m gpu`):
two ranks x B/2 samples with global loss normalisers + the engine's bucketed NCCL all-reduce must reproduce the
gradients and the AdamW update of one process running the whole batch."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import medvill_oracle as orc
from tests.util import dims_from_cfg, oracle_feats

pytestmark = pytest.mark.gpu

NAMES = ("enc.pooler.dense.weight", "enc.encoder.layer.0.intermediate.dense.weight", "enc.encoder.layer.1.attention.self.query.bias",
         "enc.txt_embeddings.word_embeddings.weight", "enc.img_embeddings.img_embeddings.weight", "mlm.predictions.bias", "itm.linear.bias")


ADAM_LR = 1e-3


def _make(cfg, params, batch, feats, sl, device, n_lab_global, b_global, precision):
    import medvill_b200 as m

    eng = m.PretrainEngine(dims_from_cfg(cfg), device, precision=precision, max_batch=8)
    eng.load_params(params)
    b = eng.make_batch(cls_tok=batch["cls_tok"][sl], input_ids=batch["input_ids"][sl], segment=batch["segment"][sl],
                       sep_tok=batch["sep_tok"][sl], mode=batch["mode"][sl], t_len=batch["t_len"][sl], region_idx=batch["region_idx"],
                       feats=feats[sl], txt_labels=batch["txt_labels"][sl], is_aligned=batch["is_aligned"][sl], seed=3, train=True,
                       n_lab_global=n_lab_global, batch_global=b_global)
    return eng, b


def _worker(rank, world, port, precision, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)      # plumbing only: carries the NCCL unique id
    cfg = orc.Cfg(**orc.TINY)
    params = orc.synth_params(cfg, seed=0)
    batch = orc.synthetic_batch(cfg, 4, seed=21, mode=orc.MODE_BAR)
    feats = oracle_feats(params, batch)
    n_glob = int((batch["txt_labels"] != -100).sum())
    sl = slice(2 * rank, 2 * rank + 2)
    eng, b = _make(cfg, params, batch, feats, sl, "cuda:%d" % rank, n_glob, 4, precision)

    def bcast(raw):
        box = [raw]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    eng.comm_init(rank, world, bcast)
    eng.zero_grads(); eng.stats_reset()
    eng.forward(b)
    eng.backward(b, allreduce=True)
    eng.comm_sync()                       # current stream waits for the bucket all-reduces issued during backward
    torch.cuda.synchronize()
    st = eng.read_stats()
    out = {n: eng.view(n, eng.grads).float().cpu().numpy().copy() for n in NAMES}
    # second pass: the optimizer step straight after backward, i.e. while the last bucket's all-reduce may still be in
    # flight (mv_adamw_step updates layers + heads behind the mid event, then the embeddings)
    before = {n: eng.view(n).float().cpu().numpy().copy() for n in NAMES}
    eng.zero_grads(); eng.stats_reset()
    eng.forward(b)
    eng.backward(b, allreduce=True)
    eng.adamw_step(lr=ADAM_LR)
    torch.cuda.synchronize()
    assert float(eng.grads.abs().max()) == 0.0                        # zero_grad fused into both ranges
    after = {n: eng.view(n).float().cpu().numpy().copy() for n in NAMES}
    q.put((rank, st, out, before, after))
    dist.destroy_process_group()


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 2e-2)])
def test_two_ranks_equal_one_process_big_batch(precision, tol):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cfg = orc.Cfg(**orc.TINY)
    params = orc.synth_params(cfg, seed=0)
    batch = orc.synthetic_batch(cfg, 4, seed=21, mode=orc.MODE_BAR)
    feats = oracle_feats(params, batch)
    n_glob = int((batch["txt_labels"] != -100).sum())
    eng, b = _make(cfg, params, batch, feats, slice(0, 4), "cuda:0", n_glob, 4, precision)
    eng.zero_grads(); eng.stats_reset()
    eng.forward(b); eng.backward(b)
    torch.cuda.synchronize()
    ref_stats = eng.read_stats()
    ref = {n: eng.view(n, eng.grads).float().cpu().numpy().copy() for n in NAMES}
    eng.close()

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, precision, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    # loss sums add up across ranks; the all-reduced gradients are identical on both ranks and equal the gradient of
    # ONE process running all 4 samples (global normalisers: no post-hoc averaging)
    assert abs(res[0][1]["mlm_loss_sum"] + res[1][1]["mlm_loss_sum"] - ref_stats["mlm_loss_sum"]) <= tol * abs(ref_stats["mlm_loss_sum"])
    assert res[0][1]["itm_correct"] + res[1][1]["itm_correct"] == ref_stats["itm_correct"]
    for n in NAMES:
        for r in range(2):
            assert np.abs(res[r][2][n] - ref[n]).max() <= tol * np.abs(ref[n]).max() + 1e-9, (n, r)
        assert np.array_equal(res[0][2][n], res[1][2][n])          # NCCL all-reduce: bitwise identical replicas
    # AdamW right behind backward: replicas stay bitwise identical, and the first step moves every element with a
    # non-negligible gradient by lr * sqrt(1-b2)/(1-b1) * m/(sqrt(v)+eps) = lr * g / (|g| + eps') ~ lr * sign(g)
    for n in NAMES:
        assert np.array_equal(res[0][4][n], res[1][4][n]), n
        g = res[0][2][n].astype(np.float64)
        want = res[0][3][n] - ADAM_LR * np.sqrt(1 - 0.999) / (1 - 0.9) * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-6)
        sure = np.abs(g) > 1e-3 * np.abs(g).max()
        assert sure.any()
        assert np.abs(res[0][4][n] - want)[sure].max() <= 0.05 * ADAM_LR, n


def _finetune_worker(rank, world, port, precision, q):
    """finetune.py:370-376 under one process per GPU: each rank runs finetune_step on its half of the fixture batch."""
    import torch.distributed as dist

    from tests.test_finetune_gpu import fixture_batch, make_model, oracle_feats as ft_feats
    from tests.util import load_golden

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, cfg = load_golden("finetune_tiny_s2s")
    batch = fixture_batch(g, cfg)
    params = orc.synth_params(cfg, seed=0)
    if rank == 1:                                   # replicas start different: init_distributed must broadcast rank 0's weights
        params = {k: (v + 0.01 if v.dtype.is_floating_point and k.startswith("mlm.") else v) for k, v in params.items()}
    feats = ft_feats(orc.synth_params(cfg, seed=0), batch)
    import medvill_b200  # noqa: F401
    model = make_model(cfg, precision, params).to("cuda:%d" % rank).train()
    assert model.init_distributed() == world
    sl = slice(2 * rank, 2 * rank + 2)
    t = lambda k: torch.as_tensor(batch[k][sl])
    out = model.finetune_step(None, t("input_ids"), t("segment_ids"), None, t("masked_ids"), t("masked_pos"), t("masked_weights"),
                              optimizer=None, mode=t("mode"), t_len=t("t_len"), feats=feats[sl])
    eng = model.engine()
    eng.comm_sync()
    torch.cuda.synchronize()
    names = [n for n in NAMES if n not in orc.FT_NO_GRAD and not n.startswith("itm.")]
    q.put((rank, out["loss"], {n: eng.view(n, eng.grads).float().cpu().numpy().copy() for n in names}))
    dist.destroy_process_group()


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 3e-2)])
def test_finetune_two_ranks_average_gradients_like_ddp(precision, tol):
    """Each rank normalises by its own masked-token count and the gradients are AVERAGED (DistributedDataParallel,
    finetune.py:376) == mean of the oracle's gradients of the two half batches."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tests.test_finetune_gpu import fixture_batch, oracle_feats as ft_feats
    from tests.util import load_golden

    g, cfg = load_golden("finetune_tiny_s2s")
    batch = fixture_batch(g, cfg)
    params = orc.synth_params(cfg, seed=0)
    feats = ft_feats(params, batch)
    halves = []
    for r in range(2):
        sl = slice(2 * r, 2 * r + 2)
        half = {k: (v[sl] if isinstance(v, np.ndarray) and v.shape[:1] == (4,) else v) for k, v in batch.items()}
        halves.append(orc.finetune_loss_and_grads(params, half, cfg, feats=feats[sl]))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_finetune_worker, args=(r, 2, port, precision, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    for r in range(2):
        assert abs(res[r][1] - halves[r]["loss"]) <= max(tol, 1e-3) * abs(halves[r]["loss"]), (r, res[r][1], halves[r]["loss"])
    for n in res[0][2]:
        want = 0.5 * (halves[0]["grads"][n] + halves[1]["grads"][n]).numpy()
        for r in range(2):
            assert np.abs(res[r][2][n] - want).max() <= tol * np.abs(want).max() + 1e-9, (n, r)
        assert np.array_equal(res[0][2][n], res[1][2][n])
