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
    q.put((rank, st, out))
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
