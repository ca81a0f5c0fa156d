"""Drop-in mirror of /root/reference/Downstream_task/Retrieval/retrieval.py:12-32 (`CXRBertForRetrieval`).

Same constructor / forward signature, same attribute tree (`.enc`, `.itm`) and therefore the same `state_dict` keys
(`enc.*`, `itm.*` — no MLM head).  All arithmetic runs in libmedvill_sm100.so through the owning `CXRBERT`'s engine
(mv_forward with no labelled rows: the MLM head is skipped); there is no PyTorch fallback.
"""
import os

import torch
import torch.nn as nn

from ..config import BertConfig
from ..models.cxrbert_origin import CXRBERT


class CXRBertForRetrieval(nn.Module):
    def __init__(self, config, args):
        super().__init__()
        if getattr(args, "weight_load", False):             # retrieval.py:17-21
            path = args.load_pretrained_model
            config = BertConfig.from_pretrained(path)
            sd = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu")
            cxrbert = CXRBERT.from_pretrained(path, state_dict=sd, config=config, args=args)
        else:                                               # retrieval.py:22-24
            if config is None:
                config = BertConfig.from_pretrained("bert-base-uncased")
            cxrbert = CXRBERT(config, args)
        self.config, self.args = config, args
        object.__setattr__(self, "_cxrbert", cxrbert)       # engine owner; NOT a registered sub-module (keeps the key set)
        self.enc = cxrbert.enc                              # retrieval.py:26-27
        self.itm = cxrbert.itm

    # nn.Module plumbing that must reach the hidden owner (its MLM-head parameters live in the same arena)
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

    def forward(self, cls_tok, input_txt, attn_mask, segment, input_img, sep_tok):
        """-> ITM logits [B, 2] fp32 (retrieval.py:29-32): pooled [CLS] of the joint encoder -> Linear(768, 2)."""
        owner = self._cxrbert
        if owner._wants_grad():                             # training: exactly the reference's composition, differentiable
            _, cls, _ = self.enc(cls_tok, input_txt, attn_mask, segment, input_img, sep_tok)
            return self.itm(cls)
        eng, batch = owner._encode(cls_tok, input_txt, attn_mask, segment, input_img, sep_tok, train=self.training)
        itm = torch.empty(batch.B, 2, dtype=torch.float32, device=eng.device)
        owner._peek_into("itm_logits", 0, itm)
        return itm
