"""GPU test of the drop-in training surface end to end: CXRDataset (JSONL + images on disk) -> DataLoader -> CXRBERT_Trainer
.train(epoch) -> .save(epoch, path) -> from_pretrained, i.e. what /root/reference/main_origin.py:118-151 drives, on a tiny
configuration; with the reference's explicit [L, L] masks and with the compact (mode, t_len) form."""
import json
import os
import types

import numpy as np
import pytest
import torch

from oracle import medvill_oracle as orc

pytestmark = pytest.mark.gpu


def make_corpus(tmp_path, n=12, img_size=128):
    from PIL import Image

    rng = np.random.RandomState(0)
    words = ["w%d" % i for i in range(200, 400)]
    rows = []
    for i in range(n):
        arr = rng.randint(0, 256, size=(img_size, img_size, 3), dtype=np.uint8)
        Image.fromarray(arr).save(tmp_path / ("img%d.png" % i))
        text = " ".join(rng.choice(words, size=rng.randint(3, 25)))
        rows.append({"id": "s%d" % i, "split": "Train", "label": "'L%d'" % (i % 3), "text": text, "img": "img%d.png" % i})
    path = tmp_path / "train.jsonl"
    path.write_text("\n".join(json.dumps(r) for r in rows))
    vocab = {"[PAD]": 0, "[UNK]": 100, "[CLS]": 101, "[SEP]": 102, "[MASK]": 103}
    vocab.update({"w%d" % i: i for i in range(200, 400)})
    used = set(vocab.values())
    vocab.update({"tok%d" % i: i for i in range(1000) if i not in used})      # 1000 entries: random_word draws ids < len(vocab)
    return str(path), vocab


@pytest.mark.parametrize("compact", [False, True])
def test_trainer_trains_saves_and_reloads(tmp_path, compact):
    import medvill_b200  # noqa: F401
    from medvill_b200.config import BertConfig
    import torchvision.transforms as T

    from medvill_b200.data import CXRDataset
    from medvill_b200.models import CXRBERT
    from medvill_b200.models.train_origin import CXRBERT_Trainer
    from medvill_b200.utils.utils import set_seed

    cfg = orc.Cfg(**orc.TINY)
    data_path, vocab = make_corpus(tmp_path, img_size=cfg.img_size)
    bc = BertConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
                    intermediate_size=cfg.inter, max_position_embeddings=cfg.max_pos, type_vocab_size=cfg.type_vocab)
    args = types.SimpleNamespace(
        bert_model="bert-base-scratch", bert_config=bc, img_hidden_sz=cfg.img_hidden, embedding_size=cfg.hidden, hidden_size=cfg.hidden,
        dropout_prob=0.1, img_postion=True, img_encoder="random-pixel", num_image_embeds=cfg.num_image_embeds, img_size=cfg.img_size,
        img_channel=3, seq_len=cfg.seq_len, max_seq_len=cfg.max_pos, Mixed=False, BAR_attn=True, attn_1d=False, disturbing_mask=False,
        s2s_prob=1.0, bi_prob=0.0, lr=2e-3, with_cuda=True, cuda_devices=None, weight_load=False, pre_trained_model_path=None,
        mlm_task=True, itm_task=True, log_freq=10, precision="bf16", max_micro_batch=4, seed=123, compact_masks=compact)
    set_seed(123)
    # get_transforms (data/helper.py:20-27) only knows the reference's 224 / 512 inputs; same ToTensor + Normalize for 128 x 128
    tf = T.Compose([T.ToTensor(), T.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])
    ds = CXRDataset(data_path, str.split, tf, args, vocab=vocab)
    sample = ds[0]
    assert len(sample) == 9 and sample[1].shape == (cfg.seq_len + 1,) and sample[2].shape == (cfg.L,)
    assert tuple(sample[3].shape) == ((2,) if compact else (cfg.L, cfg.L))
    loader = torch.utils.data.DataLoader(ds, batch_size=4, shuffle=False, num_workers=0, pin_memory=True)
    trainer = CXRBERT_Trainer(args, train_dataloader=loader, test_dataloader=loader)
    first = trainer.train(0)
    for ep in range(1, 4):
        last = trainer.train(ep)
    assert np.isfinite(first["loss"]) and np.isfinite(last["loss"])
    assert last["mlm_loss"] < first["mlm_loss"]                       # 12 steps at lr 2e-3 on 12 samples: the MLM loss must move
    assert 0.0 <= trainer.last_eval["itm_acc"] <= 100.0 and np.isfinite(trainer.last_eval["loss"])
    trainer.save(3, str(tmp_path / "out"))
    ck = tmp_path / "out" / "3"
    assert (ck / "pytorch_model.bin").is_file() and (ck / "config.json").is_file()
    sd = torch.load(str(ck / "pytorch_model.bin"), map_location="cpu")
    assert set(sd) == set(trainer.model.state_dict())
    model2 = CXRBERT.from_pretrained(str(ck), config=bc, args=args)
    for k, v in model2.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k
    # restart (train_origin.py:28-34 with --weight_load) continues Adam instead of resetting it: moments, step count and
    # the dropout counter come back from optimizer.pt
    assert (ck / "optimizer.pt").is_file()
    args2 = types.SimpleNamespace(**{**vars(args), "weight_load": True, "pre_trained_model_path": str(ck)})
    eng = trainer.model.engine()
    trainer2 = CXRBERT_Trainer(args2, train_dataloader=loader, test_dataloader=None)
    eng2 = trainer2.model.engine()
    assert eng2.step_count == eng.step_count == 12 and trainer2.model._dropout_step == trainer.model._dropout_step
    assert torch.equal(eng2.adam_m, eng.adam_m) and torch.equal(eng2.adam_v, eng.adam_v) and torch.equal(eng2.params, eng.params)
    assert float(eng2.adam_v.abs().max()) > 0
    # the same next step from both: identical update up to the order of the fp32 split-K reductions
    batch = next(iter(loader))
    before = eng.params.clone()
    outs = []
    for tr in (trainer, trainer2):
        torch.manual_seed(7)                                           # region sampling draws from the global CPU generator
        tr.model.train()                                               # trainer.train() left the first model in eval mode
        cls_tok, input_ids, txt_labels, attn_masks, img, segment, is_aligned, sep_tok, mode, t_len = tr._unpack(batch)
        outs.append(tr.model.pretrain_step(cls_tok, input_ids, txt_labels, attn_masks, img, segment, is_aligned, sep_tok, lr=args.lr,
                                           mode=mode, t_len=t_len))
    torch.cuda.synchronize()
    d1, d2 = eng.params - before, eng2.params - before
    assert abs(outs[0]["loss"] - outs[1]["loss"]) <= 1e-3 * abs(outs[0]["loss"])
    assert float((d1 - d2).norm()) <= 2e-2 * float(d1.norm()) and float(d1.norm()) > 0
