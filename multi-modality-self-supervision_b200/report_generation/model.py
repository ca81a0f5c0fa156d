"""Drop-in mirror of `BertForPreTrainingLossMask` (.../sc/pytorch_pretrained_bert/model.py:902-1054) for
tasks == 'report_generation' with img_encoding == 'fully_use_cnn' (every ResNet grid region is a visual token).

Same constructor / forward signature and the same `state_dict` keys as the reference fine-tune model
(`txt_embeddings.*`, `img_embeddings.*`, `img_encoder.model.*`, `encoder.layer.*`, `pooler.dense.*`, `cls.predictions.*`),
i.e. a pre-training checkpoint loads through the reference's own rename rule (finetune.py:338-339, `pretrain_to_finetune_key`).
All arithmetic runs in libmedvill_sm100.so through an internal `CXRBERT` engine owner; what differs from the pre-training
step is expressed through mv_batch switches (prefix [SEP] position, prefix token type, [PAD] lookup gradient), the
fine-tune mask modes, LayerNorm eps 1e-5, per-row loss weights and mv_bert_adam_step.  There is no PyTorch fallback.

`forward(...)` returns (masked_lm_loss, dummy) like the reference, as detached tensors (evaluation of the loss);
training uses `finetune_step(...)`: forward + weighted masked-LM CE + backward (+ BertAdam.step when an optimizer is passed).
"""
import types

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..config import BertConfig
from ..models.cxrbert_origin import CXRBERT


def pretrain_to_finetune_key(key):
    """finetune.py:338-339: `key.replace('enc.', '').replace('mlm.', 'cls.')`"""
    return key.replace("enc.", "").replace("mlm.", "cls.")


class BertForPreTrainingLossMask(nn.Module):
    def __init__(self, config, args, num_labels=2, len_vis_input=None, tasks="report_generation"):
        super().__init__()
        if tasks != "report_generation":
            raise NotImplementedError("only tasks='report_generation' is on the B200 path (VQA is out of scope)")
        if getattr(args, "img_encoding", "fully_use_cnn") != "fully_use_cnn":
            raise NotImplementedError("img_encoding must be 'fully_use_cnn' (all grid regions, model.py:36-54)")
        img_size = int(getattr(args, "img_size", 512))
        grid = (img_size // 32) ** 2
        if len_vis_input is None:
            len_vis_input = grid
        if len_vis_input != grid:
            raise _lib.MedvillError("len_vis_input (%d) must equal the ResNet grid (%d regions for %dx%d images)"
                                    % (len_vis_input, grid, img_size, img_size))
        self.config, self.args, self.num_labels, self.len_vis_input, self.tasks = config, args, num_labels, len_vis_input, tasks
        cfg = BertConfig(**{**config.__dict__, "layer_norm_eps": 1e-5}) if isinstance(config, BertConfig) else config
        if not isinstance(config, BertConfig):
            cfg.layer_norm_eps = 1e-5                                    # BertLayerNorm(eps=1e-5) everywhere (model.py:238,327,367)
        max_b = int(getattr(args, "max_len_b", 0)) or (cfg.max_position_embeddings - len_vis_input - 3)
        drop = float(getattr(cfg, "hidden_dropout_prob", 0.1))
        inner = types.SimpleNamespace(img_hidden_sz=args.img_hidden_sz, embedding_size=args.hidden_size, hidden_size=args.hidden_size,
                                      dropout_prob=drop, img_encoder="random-pixel", num_image_embeds=len_vis_input, img_size=img_size,
                                      seq_len=max_b, lr=getattr(args, "learning_rate", 3e-5), precision=getattr(args, "precision", "bf16"),
                                      max_micro_batch=int(getattr(args, "max_micro_batch", 64)), seed=int(getattr(args, "seed", 123)),
                                      allow_random_trunk=bool(getattr(args, "allow_random_trunk", False)),
                                      resnet_weights=getattr(args, "resnet_weights", None), engine_flags=int(getattr(args, "engine_flags", 0)))
        cx = CXRBERT(cfg, inner)
        object.__setattr__(self, "_cxrbert", cx)                          # engine owner, not a registered sub-module
        self.txt_embeddings = cx.enc.txt_embeddings                        # model.py:907-923
        self.img_embeddings = cx.enc.img_embeddings
        self.img_encoder = cx.enc.img_encoder
        self.encoder = cx.enc.encoder
        self.pooler = cx.enc.pooler
        self.cls = cx.mlm
        self.A = len_vis_input + 2
        self.L = self.A + max_b + 1

    # nn.Module plumbing reaches the hidden owner (its parameters are the same objects)
    def _apply(self, fn, *a, **k):
        self._cxrbert._apply(fn, *a, **k)
        return self

    def train(self, mode=True):
        self._cxrbert.train(mode)
        return super().train(mode)

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = {k: v for k, v in state_dict.items() if not k.endswith("position_ids")}
        out = super().load_state_dict(sd, strict=strict, **kw)
        self._cxrbert.enc.img_encoder._exec = None
        self._cxrbert.sync_params()
        return out

    def engine(self, min_batch=1):
        return self._cxrbert.engine(min_batch)

    def init_distributed(self):
        """finetune.py:370-376 wraps the model in DistributedDataParallel; here one call joins the engine's NCCL
        communicator (one process per GPU) and `finetune_step` averages the ranks' gradients as DDP does."""
        return self._cxrbert.init_distributed()

    # -- batch assembly ------------------------------------------------------------------------------------------------
    def _labelled_rows(self, masked_pos, masked_ids, masked_weights):
        """(lab_rows, lab_labels, lab_weights, sum_w): one row per DISTINCT masked position with weight > 0; a position
        listed twice (the forced final-[SEP] mask can repeat a sampled one, data_loader.py:368-372) carries weight 2."""
        pos = torch.as_tensor(masked_pos).cpu().numpy().astype(np.int64)
        ids = torch.as_tensor(masked_ids).cpu().numpy().astype(np.int64)
        w = torch.as_tensor(masked_weights).cpu().numpy().astype(np.float64)
        B = pos.shape[0]
        flat = (np.arange(B)[:, None] * self.L + pos)[w > 0]
        rows, inv = np.unique(flat, return_inverse=True)
        weights = np.bincount(inv, weights=w[w > 0], minlength=rows.size)
        labels = np.zeros(rows.size, dtype=np.int64)
        labels[inv] = ids[w > 0]
        return torch.from_numpy(rows), torch.from_numpy(labels), torch.from_numpy(weights.astype(np.float32)), float(w.sum())

    def _run(self, img, input_ids, token_type_ids, attention_mask, masked_lm_labels, masked_pos, masked_weights, train, mode=None,
             t_len=None, feats=None, backward=False, drop_worst_ratio=0.0):
        cx, A = self._cxrbert, self.A
        input_ids = torch.as_tensor(input_ids)
        token_type_ids = torch.as_tensor(token_type_ids)
        if tuple(input_ids.shape[1:]) != (self.L,):
            raise _lib.MedvillError("input_ids must be [B, %d]" % self.L)
        B = int(input_ids.shape[0])
        eng = cx.engine(min(B, int(cx.args.max_micro_batch)))
        if mode is None:
            am = torch.as_tensor(attention_mask)
            if am.dim() == 2 and am.shape[1] == 2:                       # compact (mode, t_len) from Preprocess4Seq2seq
                mode, t_len = am[:, 0].to(torch.uint8), am[:, 1].to(torch.int32)
            else:
                mode, t_len = cx.classify_mask(am, eng)
        prefix = token_type_ids[:, :A]
        ptype = int(prefix[0, 0])
        if not bool((prefix == ptype).all()):
            raise _lib.MedvillError("token types of [CLS] / regions / [SEP] must be one value per batch (0, or 4 with new_segment_ids)")
        rows, labels, weights, sum_w = self._labelled_rows(masked_pos, masked_lm_labels, masked_weights)
        denom = sum_w + 1e-5                                             # model.py:1009 with drop_worst_ratio = 0
        eng.stats_reset()
        cap = eng.max_batch
        chunks = [(s, min(B, s + cap)) for s in range(0, B, cap)]
        keep = 0
        if drop_worst_ratio > 0:
            # model.py:1006-1010 ranks the samples of the whole batch: it must be one micro-batch; the kept set and its
            # denominator live on the device and mlm_loss_sum arrives normalised (mv_batch.drop_worst_keep)
            keep = int(B * (1 - drop_worst_ratio))                       # :1007, the reference's own rounding
            if len(chunks) > 1:
                raise _lib.MedvillError("drop_worst_ratio > 0 needs the batch (%d) in one micro-batch: raise args.max_micro_batch (%d)"
                                        % (B, cap))
            if keep < 1:
                raise _lib.MedvillError("drop_worst_ratio %.3f keeps no sample of a batch of %d" % (drop_worst_ratio, B))
            denom = 1.0
        # DistributedDataParallel semantics (finetune.py:376): every rank normalises by its OWN denominator and the gradients
        # are averaged; folded into the loss scale so that the engine's SUM all-reduce needs no post-hoc division
        world = eng.world
        for ci, (s, e) in enumerate(chunks):
            sel = (rows >= s * self.L) & (rows < e * self.L)
            _, batch = cx._encode(input_ids[s:e, :1], input_ids[s:e, A:], None, token_type_ids[s:e, A:],
                                  None if img is None else img[s:e], input_ids[s:e, A - 1:A], train=train, mode=mode[s:e], t_len=t_len[s:e],
                                  feats=None if feats is None else feats[s:e], lab_rows=rows[sel] - s * self.L, lab_labels=labels[sel],
                                  lab_weights=weights[sel], n_lab_global=denom * world, batch_global=float(B), sep_position=A - 1,
                                  prefix_type=ptype, pad_lookup_grad=True, drop_worst_keep=keep)
            if backward:
                eng.backward(batch, allreduce=(world > 1 and ci == len(chunks) - 1))
        return eng, denom

    def forward(self, img, _, input_ids, token_type_ids=None, attention_mask=None, masked_lm_labels=None, ans_labels=None,
                masked_pos=None, masked_weights=None, task_idx=None, drop_worst_ratio=0.2, vqa_inference=False, ans_type=None,
                mode=None, t_len=None, feats=None):
        """-> (masked_lm_loss, dummy) as at model.py:968-1054, including Luo's drop-worst (:1003-1010; finetune.py passes
        --max_drop_worst_ratio, default 0, after --drop_after epochs, :179-180,440)."""
        if vqa_inference or ans_labels is not None:
            raise NotImplementedError("VQA is out of scope of the B200 path")
        eng, denom = self._run(img, input_ids, token_type_ids, attention_mask, masked_lm_labels, masked_pos, masked_weights,
                               train=self.training, mode=mode, t_len=t_len, feats=feats, drop_worst_ratio=drop_worst_ratio)
        st = eng.read_stats()
        dev = eng.device
        return torch.tensor(st["mlm_loss_sum"] / denom, device=dev), torch.zeros(1, device=dev)

    def finetune_step(self, img, input_ids, token_type_ids, attention_mask, masked_lm_labels, masked_pos, masked_weights,
                      optimizer=None, mode=None, t_len=None, feats=None, lazy=False, drop_worst_ratio=0.0):
        """One training step of finetune.py:427-463: forward, masked-LM loss, backward, optimizer.step() + zero_grad()."""
        eng, denom = self._run(img, input_ids, token_type_ids, attention_mask, masked_lm_labels, masked_pos, masked_weights,
                               train=True, mode=mode, t_len=t_len, feats=feats, backward=True, drop_worst_ratio=drop_worst_ratio)
        n_masked = float(torch.as_tensor(masked_weights).sum())
        if optimizer is not None:
            if getattr(optimizer, "_engine", None) is None:
                optimizer._engine = eng
            optimizer.step()
            optimizer.zero_grad()
        finish = lambda st: dict(loss=st["mlm_loss_sum"] / denom, mlm_correct=st["mlm_correct"], n_masked=n_masked)
        if lazy:
            pending = eng.read_stats_async()
            return lambda: finish(pending())
        return finish(eng.read_stats())
