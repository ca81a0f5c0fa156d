"""CPU: host-side mirror of the reference interface (state_dict contract, synthetic data) and the data-parallel
normalisation / sharding logic under a 2-rank gloo group."""
import os
import types

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import medvill_b200 as m
from medvill_b200.config import BertConfig
from medvill_b200.data.synthetic import synthetic_batch
from medvill_b200.models import CXRBERT


def base_args(**kw):
    a = types.SimpleNamespace(img_hidden_sz=2048, embedding_size=768, hidden_size=768, dropout_prob=0.1, img_encoder="random-pixel",
                              num_image_embeds=180, img_size=512, seq_len=253, lr=1e-5)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def test_state_dict_contract():
    model = CXRBERT(BertConfig.from_pretrained("bert-base-uncased"), base_args())
    sd = model.state_dict()
    assert len(sd) == 531
    assert sum(p.numel() for p in model.parameters()) == 135188092
    assert sum(p.numel() for p in model.parameters() if p.requires_grad) == 111680060
    assert sum(1 for k in sd if k.startswith("enc.img_encoder.model.")) == 318
    for k in ("enc.txt_embeddings.word_embeddings.weight", "enc.img_embeddings.img_embeddings.weight",
              "enc.img_embeddings.LayerNorm.weight", "enc.encoder.layer.11.attention.self.query.bias",
              "enc.encoder.layer.0.output.LayerNorm.bias", "enc.pooler.dense.weight", "mlm.predictions.bias",
              "mlm.predictions.transform.LayerNorm.weight", "mlm.predictions.decoder.weight", "itm.linear.weight"):
        assert k in sd, k
    assert sd["mlm.predictions.decoder.weight"].data_ptr() == sd["enc.txt_embeddings.word_embeddings.weight"].data_ptr()
    assert sd["enc.img_embeddings.img_embeddings.weight"].shape == (768, 2048)
    assert set(m.param_map(model.dims(), m.query_layout(model.dims()))) == {n for n, p in model.named_parameters() if p.requires_grad}
    # a transformers-3.x checkpoint carries a position_ids buffer: accepted and ignored
    sd2 = dict(sd)
    sd2["enc.txt_embeddings.position_ids"] = torch.arange(512)[None]
    model.load_state_dict(sd2)


def test_synthetic_batch_semantics():
    b = synthetic_batch(6, seed=7)
    A, T, L = 182, 254, 436
    assert b["input_ids"].shape == (6, T) and b["txt_labels"].shape == (6, L) and b["image"].shape == (6, 3, 512, 512)
    for i in range(6):
        t = int(b["t_len"][i])
        ids, lab = b["input_ids"][i], b["txt_labels"][i]
        assert int(ids[t - 1]) == 102 and (ids[t:] == 0).all()          # [SEP] then [PAD]
        assert (lab[:A] == -100).all() and (lab[A + t - 1:] == -100).all()
        sel = lab != -100
        assert 1 <= int(sel.sum()) <= t - 1
        assert int(b["mode"][i]) == 2


def _ddp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank holds B samples; the loss normalisers must be GLOBAL so that SUM-all-reduced gradients equal the
    # gradient of one process running the concatenated batch
    b = synthetic_batch(4, seed=100 + rank, img_size=32)
    n_local = int((b["txt_labels"] != -100).sum())
    cnt = torch.tensor([float(n_local), 4.0])
    dist.all_reduce(cnt)
    token_loss = torch.rand(n_local, generator=torch.Generator().manual_seed(rank))       # stand-in per-token CE
    sample_loss = torch.rand(4, generator=torch.Generator().manual_seed(10 + rank))
    contrib = torch.tensor([float(token_loss.sum() / cnt[0] + sample_loss.sum() / cnt[1])])
    dist.all_reduce(contrib)
    q.put((rank, n_local, float(cnt[0]), float(contrib), token_loss.tolist(), sample_loss.tolist()))
    dist.destroy_process_group()


def test_two_rank_global_normalisation_equals_single_process_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    n_total = res[0][1] + res[1][1]
    assert res[0][2] == res[1][2] == n_total
    tok = np.concatenate([res[0][4], res[1][4]])
    smp = np.concatenate([res[0][5], res[1][5]])
    single = tok.mean() + smp.mean()                                    # one process, batch of 8: CE means
    assert abs(res[0][3] - single) < 1e-5 and abs(res[1][3] - single) < 1e-5


def test_mask_predicates_exhaustive_host_build():
    """csrc/tests/mask_test (host-only build of mask.cuh): per-row intervals and the tile-skipping predicates of all six mask
    modes (four pre-training, two fine-tune) against brute force over every (A, t_len) of small layouts"""
    import subprocess

    from medvill_b200 import _lib

    exe = os.path.join(_lib.CSRC, "tests", "mask_test")
    out = subprocess.run(["make", "-C", _lib.CSRC, "tests/mask_test"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "MASK TEST PASSED" in run.stdout, run.stdout[-2000:]


def test_main_origin_parser_keeps_the_reference_command_line():
    """every flag of the reference's main_origin.py:80-139 (fixture: oracle/make_golden_cli.py) exists with the same default;
    `choices` are equal or a superset (the hard-coded checkpoint paths / model names of the reference are not enforced);
    --output_path defaults to None here because the reference creates `output/<now>` at import time"""
    import json

    from medvill_b200.main_origin import build_parser

    here = os.path.dirname(os.path.abspath(__file__))
    ref = json.load(open(os.path.join(here, "golden", "main_origin_flags.json")))
    ours = {a.option_strings[0]: a for a in build_parser()._actions if a.option_strings and a.option_strings[0] != "-h"}
    assert len(ref) == 42
    for flag, spec in ref.items():
        assert flag in ours, flag
        act = ours[flag]
        if flag != "--output_path":
            assert act.default == spec.get("default"), (flag, act.default, spec.get("default"))
        if "choices" in spec and act.choices is not None:
            assert set(spec["choices"]) <= set(act.choices), flag
        if spec.get("type") in ("int", "float", "str"):
            assert act.type is {"int": int, "float": float, "str": str}[spec["type"]], flag
    assert set(ours) - set(ref) == {"--precision", "--compact_masks", "--max_micro_batch", "--allow_random_trunk", "--resnet_weights"}
