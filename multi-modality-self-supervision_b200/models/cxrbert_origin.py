"""Drop-in mirror of /root/reference/models/cxrbert_origin.py for the pre-training path.

Same class names, constructor / forward signatures, attribute tree and `state_dict` keys as the reference
(`CXRBERT`, `CXRBertEncoder`, `ImageBertEmbeddings`, `BertPreTrainingHeads`, `ImageTextMatching`), but the
modules are *parameter containers*: every trainable tensor is a view into the flat arena of a `PretrainEngine` and all
arithmetic runs in libmedvill_sm100.so (tcgen05 GEMMs, fused masked attention, fused LN / embedding / CE kernels).
There is no PyTorch fallback: on a CPU device `forward` raises.

What differs from the reference, deliberately:
  * the attention mask tensor is accepted for compatibility, but is reduced on the device to (mode, t_len) per sample
    and *validated* against the closed-form predicate; arbitrary masks are rejected loudly;
  * `CXRBertEncoder.forward` returns `None` for the attentions slot (the reference keeps 7 GB of them and drops them);
  * training uses `CXRBERT.pretrain_step` (fused forward + losses + backward + all-reduce + AdamW); `forward` returns
    detached tensors.
"""
import json
import os

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..config import BertConfig
from ..engine import EngineDims, PretrainEngine
from ..utils.utils import deterministic_requested
from .image import ImageEncoder_cnn

_FUSED_MSG = ("%s.forward is fused into libmedvill_sm100 (mv_forward); call CXRBERT / CXRBertEncoder.forward or "
              "CXRBERT.pretrain_step instead")


class _JointForward(torch.autograd.Function):
    """CXRBERT.forward / CXRBertEncoder.forward under torch.autograd: the reference's own training loop
    (`mlm, itm = model(...)`; `loss = CE(mlm.transpose(1, 2), labels) + CE(itm, is_aligned)`; `loss.backward()`;
    `optimizer.step()` — models/train_origin.py:106-131) and Retrieval/retrieval.py:29-32 run on it unmodified.
    forward = mv_forward (+ mv_full_logits); backward = mv_backward_external, which accumulates straight into the
    gradient arena the nn.Parameters' .grad tensors are views of (so no per-parameter gradient is returned here)."""

    @staticmethod
    def forward(ctx, anchor, owner, kind, cls_tok, input_txt, attn_mask, segment, input_img, sep_tok):
        if int(input_txt.shape[0]) > owner.engine().max_batch:
            raise _lib.MedvillError("autograd forward: batch %d exceeds args.max_micro_batch = %d (one set of saved activations)"
                                    % (int(input_txt.shape[0]), owner.engine().max_batch))
        eng, batch = owner._encode(cls_tok, input_txt, attn_mask, segment, input_img, sep_tok, train=True)
        owner._generation += 1
        ctx.owner, ctx.eng, ctx.batch, ctx.kind, ctx.generation = owner, eng, batch, kind, owner._generation
        d = eng.dims
        if kind == "logits":
            logits = eng.full_logits(batch)
            itm = torch.empty(batch.B, 2, dtype=torch.float32, device=eng.device)
            owner._peek_into("itm_logits", 0, itm)
            return logits, itm
        seq = torch.empty(batch.B * d.L, d.hidden, dtype=eng.act_dtype, device=eng.device)
        owner._peek_into("x", d.layers, seq)
        pooled = torch.empty(batch.B, d.hidden, dtype=eng.act_dtype, device=eng.device)
        owner._peek_into("pooled", 0, pooled)
        return seq.view(batch.B, d.L, d.hidden).float(), pooled.float()

    @staticmethod
    def backward(ctx, g_a, g_b):
        owner, eng, batch = ctx.owner, ctx.eng, ctx.batch
        if owner._generation != ctx.generation or owner._engine is not eng:
            raise _lib.MedvillError("backward through a CXRBERT forward whose activations were overwritten by a later forward "
                                    "(the engine keeps ONE set of saved activations): call backward before the next forward")
        owner._prepare_grads()
        d = eng.dims
        kw = {}
        if ctx.kind == "logits":
            if g_a is not None:
                M, V, ld = batch.B * d.L, d.vocab, eng.layout["vocab_padded"]
                g = g_a.reshape(M, V)
                rows = torch.nonzero((g != 0).any(dim=1)).reshape(-1)       # CE(ignore_index=-100): unlabelled rows are exactly 0
                if rows.numel():
                    dl = torch.zeros(rows.numel(), ld, dtype=eng.act_dtype, device=eng.device)
                    dl[:, :V] = g.index_select(0, rows)
                    kw.update(rows=rows, dlogits=dl)
            if g_b is not None:
                kw["d_itm"] = g_b.to(torch.float32).contiguous()
        else:
            if g_a is not None:
                kw["d_seq"] = g_a.reshape(batch.B * d.L, d.hidden).to(eng.act_dtype).contiguous()
            if g_b is not None:
                kw["d_pooled"] = g_b.to(eng.act_dtype).contiguous()
        eng.backward_external(batch, allreduce=eng.world > 1, **kw)
        if eng.world > 1:            # DistributedDataParallel semantics: the ranks' gradients are averaged
            eng.comm_sync()
            eng.grads.mul_(1.0 / eng.world)
        owner._generation += 1       # the saved activations are consumed
        return (None,) * 9


# ---- parameter containers with the upstream (HF BertModel) attribute names -------------------------------------------
class _Embeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=getattr(config, "pad_token_id", 0))
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, *a, **k):
        raise RuntimeError(_FUSED_MSG % "txt_embeddings")


class ImageBertEmbeddings(nn.Module):
    """reference: models/cxrbert_origin.py:12-35 (shares LayerNorm / position / type tables with the text embeddings)"""

    def __init__(self, args, embeddings):
        super().__init__()
        self.args = args
        self.img_embeddings = nn.Linear(args.img_hidden_sz, args.embedding_size)
        self.token_type_embeddings = embeddings.token_type_embeddings
        self.LayerNorm = embeddings.LayerNorm
        self.dropout = nn.Dropout(args.dropout_prob)
        self.position_embeddings = embeddings.position_embeddings

    def forward(self, input_imgs, img_pos, token_type_ids):
        raise RuntimeError(_FUSED_MSG % "ImageBertEmbeddings")


class _Dense(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.dense = nn.Linear(i, o)


class _DenseLN(nn.Module):
    def __init__(self, i, o, eps):
        super().__init__()
        self.dense = nn.Linear(i, o)
        self.LayerNorm = nn.LayerNorm(o, eps=eps)


class _SelfAttention(nn.Module):
    def __init__(self, H):
        super().__init__()
        self.query, self.key, self.value = nn.Linear(H, H), nn.Linear(H, H), nn.Linear(H, H)


class _Attention(nn.Module):
    def __init__(self, H, eps):
        super().__init__()
        self.self = _SelfAttention(H)
        self.output = _DenseLN(H, H, eps)


class _Layer(nn.Module):
    def __init__(self, config):
        super().__init__()
        H, I = config.hidden_size, config.intermediate_size
        self.attention = _Attention(H, config.layer_norm_eps)
        self.intermediate = _Dense(H, I)
        self.output = _DenseLN(I, H, config.layer_norm_eps)


class _Encoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(config) for _ in range(config.num_hidden_layers)])


def _init_bert_weights(module, std):
    """upstream BertPreTrainedModel._init_weights: N(0, 0.02) linears/embeddings, zero biases, unit LayerNorm."""
    for m in module.modules():
        if isinstance(m, nn.Linear):
            m.weight.data.normal_(mean=0.0, std=std)
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.Embedding):
            m.weight.data.normal_(mean=0.0, std=std)
            if m.padding_idx is not None:
                m.weight.data[m.padding_idx].zero_()
        elif isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)


class CXRBertEncoder(nn.Module):
    """reference: models/cxrbert_origin.py:37-130"""

    def __init__(self, config, args):
        super().__init__()
        self.config, self.args = config, args
        self.txt_embeddings = _Embeddings(config)
        self.img_embeddings = ImageBertEmbeddings(args, self.txt_embeddings)
        if getattr(args, "img_encoder", "random-pixel") == "ViT":
            raise NotImplementedError("the ViT image path is broken in the reference (models/image.py:105-110) and out of scope")
        self.img_encoder = ImageEncoder_cnn(args)
        for p in self.img_encoder.parameters():     # entire trunk frozen (cxrbert_origin.py:66-70)
            p.requires_grad = False
        self.encoder = _Encoder(config)
        self.pooler = _Dense(config.hidden_size, config.hidden_size)
        std = getattr(config, "initializer_range", 0.02)
        for m in (self.txt_embeddings, self.encoder, self.pooler):
            _init_bert_weights(m, std)
        self._owner = None   # set by CXRBERT: the engine lives on the top-level module

    def get_extended_attn_mask(self, attn_mask):
        """models/cxrbert_origin.py:75-85 — kept for API compatibility; the CUDA path never materialises it."""
        if attn_mask.dim() == 2:
            ext = attn_mask.unsqueeze(1).unsqueeze(2)
        elif attn_mask.dim() == 3:
            ext = attn_mask.unsqueeze(1)
        else:
            raise NotImplementedError
        return (1.0 - ext.to(dtype=torch.float16)) * -10000.0

    def forward(self, cls_tok, input_txt, attn_mask, segment, input_img, sep_tok):
        if self._owner is None:
            raise RuntimeError("CXRBertEncoder must be owned by a CXRBERT (the engine lives on the top-level module)")
        owner = self._owner()
        if owner._wants_grad():
            seq, pooled = _JointForward.apply(owner._anchor(), owner, "enc", cls_tok, input_txt, attn_mask, segment, input_img, sep_tok)
            return seq, pooled, None
        eng, batch = owner._encode(cls_tok, input_txt, attn_mask, segment, input_img, sep_tok, train=self.training)
        d = eng.dims
        seq = torch.empty(batch.B * d.L, d.hidden, dtype=eng.act_dtype, device=eng.device)
        owner._peek_into("x", d.layers, seq)
        pooled = torch.empty(batch.B, d.hidden, dtype=eng.act_dtype, device=eng.device)
        owner._peek_into("pooled", 0, pooled)
        return seq.view(batch.B, d.L, d.hidden).float(), pooled.float(), None


class _ItmHead(torch.autograd.Function):
    """Linear(hidden, 2) on the pooled output through mv_itm_head (forward and backward)"""

    @staticmethod
    def forward(ctx, x, weight, bias):
        if not x.is_cuda:
            raise _lib.MedvillError("ImageTextMatching needs CUDA tensors (sm_100a): there is no CPU fallback")
        act = torch.float32 if x.dtype == torch.float32 else torch.bfloat16
        xa = x.detach().to(act).contiguous()
        B, H = xa.shape
        out = torch.empty(B, 2, dtype=torch.float32, device=x.device)
        prec = _lib.MV_PREC_FP32 if act == torch.float32 else _lib.MV_PREC_BF16
        _lib.check(_lib.lib().mv_itm_head(_lib.ptr(xa), _lib.ptr(weight.detach()), _lib.ptr(bias.detach()), _lib.ptr(out), B, H, None, None,
                                          None, None, prec, _lib.stream_ptr(x.device)), "mv_itm_head")
        ctx.save_for_backward(xa, weight, bias)
        ctx.prec, ctx.x_dtype = prec, x.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        xa, weight, bias = ctx.saved_tensors
        B, H = xa.shape
        g = g.to(torch.float32).contiguous()
        dx = torch.empty_like(xa)
        dw, db = torch.zeros(2, H, dtype=torch.float32, device=xa.device), torch.zeros(2, dtype=torch.float32, device=xa.device)
        scratch = torch.empty(B, 2, dtype=torch.float32, device=xa.device)
        _lib.check(_lib.lib().mv_itm_head(_lib.ptr(xa), _lib.ptr(weight.detach()), _lib.ptr(bias.detach()), _lib.ptr(scratch), B, H,
                                          _lib.ptr(g), _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), ctx.prec, _lib.stream_ptr(xa.device)),
                   "mv_itm_head")
        return dx.to(ctx.x_dtype), dw, db


class ImageTextMatching(nn.Module):
    """reference: models/cxrbert_origin.py:164-173"""

    def __init__(self, hidden):
        super().__init__()
        self.linear = nn.Linear(hidden, 2)

    def forward(self, x):
        """Linear(hidden, 2) on a pooled output, as Downstream_task/Retrieval/retrieval.py:29-32 composes it"""
        return _ItmHead.apply(x, self.linear.weight, self.linear.bias)


class BertLayerNorm(nn.Module):
    """TF-style LayerNorm container, eps inside the sqrt (models/cxrbert_origin.py:189-202)"""

    def __init__(self, hidden_size, eps=1e-5):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.bias = nn.Parameter(torch.zeros(hidden_size))
        self.variance_epsilon = eps


class BertPredictionHeadTransform(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=1e-5)


class BertLMPredictionHead(nn.Module):
    """decoder weight IS the word-embedding Parameter (models/cxrbert_origin.py:225-233)"""

    def __init__(self, config, bert_model_embedding_weights):
        super().__init__()
        self.transform = BertPredictionHeadTransform(config)
        self.decoder = nn.Linear(bert_model_embedding_weights.size(1), bert_model_embedding_weights.size(0), bias=False)
        self.decoder.weight = bert_model_embedding_weights
        self.bias = nn.Parameter(torch.zeros(bert_model_embedding_weights.size(0)))


class BertPreTrainingHeads(nn.Module):
    def __init__(self, config, bert_model_embedding_weights):
        super().__init__()
        self.predictions = BertLMPredictionHead(config, bert_model_embedding_weights)

    def forward(self, sequence_output):
        raise RuntimeError(_FUSED_MSG % "BertPreTrainingHeads")


# ---- the model ---------------------------------------------------------------------------------------------------------
class CXRBERT(nn.Module):
    """Multimodal BERT: Masked Language Model + Image Text Matching (models/cxrbert_origin.py:132-149)."""

    def __init__(self, config, args):
        super().__init__()
        import weakref

        self.config, self.args = config, args
        self.enc = CXRBertEncoder(config, args)
        self.mlm = BertPreTrainingHeads(config, self.enc.txt_embeddings.word_embeddings.weight)
        self.itm = ImageTextMatching(args.hidden_size)
        self.enc._owner = weakref.ref(self)
        self._engine = None
        self._dropout_step = 0
        self._generation = 0         # bumped by every autograd forward / backward (one set of saved activations)

    # -- torch.autograd drop-in path --
    def _wants_grad(self):
        return self.training and torch.is_grad_enabled() and self.enc.pooler.dense.weight.requires_grad

    def _anchor(self):
        """a leaf that requires grad, so that autograd calls _JointForward.backward (its own gradient is None there: the
        engine accumulates into the arena that IS every parameter's .grad)"""
        eng = self.engine()
        eng.refresh_shadow()         # an external optimizer may have stepped the fp32 master parameters in place
        return self.itm.linear.bias

    def _prepare_grads(self):
        """Before an accumulating backward: parameters whose .grad was dropped (optimizer.zero_grad(set_to_none=True)) mean a
        fresh gradient -> zero the arena once and re-attach the .grad views (in-place zero_grad keeps the views)."""
        eng = self._engine
        named = self._trainable()
        if all(p.grad is None for p in named.values()):
            eng.zero_grads()
            for n, p in named.items():
                p.grad = eng.view(n, eng.grads)
            return
        for n, p in named.items():          # mixed state (e.g. ImageTextMatching's own backward already ran): tensor by tensor
            view = eng.view(n, eng.grads)
            if p.grad is None:
                view.zero_()
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view

    # -- engine plumbing --
    def dims(self):
        c, a = self.config, self.args
        if getattr(c, "hidden_act", "gelu") != "gelu":
            raise _lib.MedvillError("only the exact-erf GELU of the reference is implemented")
        grid = (a.img_size // 32) ** 2
        return EngineDims(hidden=c.hidden_size, heads=c.num_attention_heads, layers=c.num_hidden_layers, inter=c.intermediate_size,
                          vocab=c.vocab_size, max_pos=c.max_position_embeddings, type_vocab=c.type_vocab_size,
                          num_image_embeds=a.num_image_embeds, seq_len=a.seq_len, img_hidden=a.img_hidden_sz, grid=grid,
                          ln_eps=c.layer_norm_eps, head_ln_eps=1e-5,
                          # dropout sites as in the reference: BERT's own probabilities inside the encoder and the text
                          # embeddings, args.dropout_prob only for the image embeddings (cxrbert_origin.py:19)
                          dropout_p=float(getattr(c, "hidden_dropout_prob", a.dropout_prob)),
                          attn_dropout_p=float(getattr(c, "attention_probs_dropout_prob", a.dropout_prob)),
                          img_dropout_p=float(a.dropout_prob),
                          flags=int(getattr(a, "engine_flags", 0)) | (_lib.FLAG_DETERMINISTIC if deterministic_requested() else 0))

    def _trainable(self):
        """reference-named trainable parameters (aliases de-duplicated, ResNet excluded)"""
        return {n: p for n, p in self.named_parameters() if not n.startswith("enc.img_encoder.")}

    def _release_engine(self):
        if self._engine is not None:
            for p in self._trainable().values():
                p.data = p.data.clone()
                p.grad = None
            self._engine.close()
            self._engine = None

    def _apply(self, fn, *a, **k):
        # .to() / .cuda() / .float() re-allocate parameters: re-adopt lazily afterwards.  The optimizer state moves with the
        # model (host copy -> new engine); a live communicator cannot be carried over, so that case fails loudly.
        carry = None
        if self._engine is not None:
            if self._engine.world > 1:
                raise _lib.MedvillError("CXRBERT.to()/.cuda() after init_distributed(): move the model first, then initialise "
                                        "the data-parallel group (the NCCL communicator lives in the engine)")
            if self._engine.step_count > 0:
                carry = self._engine.optimizer_state_dict()
        self._release_engine()
        out = super()._apply(fn, *a, **k)
        if carry is not None:
            self._carry_optimizer = carry
        return out

    def engine(self, min_batch=1):
        dev = self.enc.pooler.dense.weight.device
        if dev.type != "cuda":
            raise _lib.MedvillError("CXRBERT needs its parameters on a CUDA device (sm_100a): there is no CPU fallback; "
                                    "call .to('cuda') first")
        cap = max(min_batch, int(getattr(self.args, "max_micro_batch", 64)))
        if self._engine is not None and self._engine.max_batch < min_batch:
            # a rebuilt engine starts with zero Adam moments, step 0 and no NCCL communicator: never do that behind the
            # caller's back once training state exists (forward / pretrain_step / eval_step chunk by max_batch instead)
            if self._engine.step_count > 0 or self._engine.world > 1:
                raise _lib.MedvillError("batch of %d exceeds the engine's micro-batch capacity %d after training started: set "
                                        "args.max_micro_batch before the first step (optimizer state and the communicator "
                                        "live in the engine)" % (min_batch, self._engine.max_batch))
            self._release_engine()
        if self._engine is None:
            eng = PretrainEngine(self.dims(), dev, precision=getattr(self.args, "precision", "bf16"), max_batch=cap)
            named = self._trainable()
            missing = set(eng.pmap) - set(named)
            if missing:
                raise _lib.MedvillError("parameters missing from the module tree: %s" % sorted(missing)[:4])
            with torch.no_grad():
                for n, p in named.items():
                    v = eng.view(n)
                    v.copy_(p.data.to(torch.float32))
                    p.data = v                      # the nn.Parameter now IS the arena slice
                    p.grad = eng.view(n, eng.grads)
            eng.refresh_shadow()
            self.enc.img_encoder.model.to(memory_format=torch.channels_last)
            if getattr(self, "_carry_optimizer", None) is not None:
                eng.load_optimizer_state_dict(self._carry_optimizer)
                self._carry_optimizer = None
            self._engine = eng
        return self._engine

    def init_distributed(self):
        """One process per GPU under an initialised torch.distributed group: hand rank 0's NCCL unique id to the engine (it
        all-reduces gradient buckets on its own stream while backward runs) and make the replicas identical — rank 0's
        weights win, as nn.DataParallel broadcast them every step (models/train_origin.py:53-55).  Returns the world size."""
        dist = torch.distributed
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return 1
        eng = self.engine()
        if eng.world == 1:
            def bcast(raw):
                box = [raw]
                dist.broadcast_object_list(box, src=0)
                return box[0]
            eng.comm_init(dist.get_rank(), dist.get_world_size(), bcast)
            # rank 0's weights win — the trainable arena AND the frozen ResNet trunk with its BatchNorm buffers (main_origin
            # seeds with seed + rank before the model is built, so without this every rank would train against its own trunk)
            trunk = [t for t in list(self.enc.img_encoder.parameters()) + list(self.enc.img_encoder.buffers())]
            for t in [eng.params] + trunk:
                if dist.get_backend() == "nccl":
                    dist.broadcast(t.data, src=0)
                else:                               # a CPU-only group (gloo) carries the id; weights go through the host
                    host = t.data.cpu()
                    dist.broadcast(host, src=0)
                    t.data.copy_(host)
            self.enc.img_encoder._exec = None       # cached dtype / layout copies of the conv weights are stale now
            self.sync_params()
        return eng.world

    def sync_params(self):
        """Call after mutating parameters from Python (e.g. load_state_dict): refreshes the bf16 GEMM-operand shadow."""
        if self._engine is not None:
            self._engine.refresh_shadow()

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = {k: v for k, v in state_dict.items() if not k.endswith("position_ids")}   # transformers-3.x buffer, ignored
        out = super().load_state_dict(sd, strict=strict, **kw)
        self.enc.img_encoder._exec = None
        self.sync_params()
        return out

    def _peek_into(self, name, layer, dst):
        import ctypes as C

        eng = self._engine
        got = C.c_int64(0)
        _lib.check(_lib.lib().mv_peek(eng._h, name.encode(), layer, _lib.ptr(dst), dst.numel() * dst.element_size(), C.byref(got),
                                      _lib.stream_ptr(eng.device)), "mv_peek")
        return dst

    def grid_features(self, input_img, eng):
        """ResNet-50 trunk on cuDNN: frozen, train-mode BN when the module is training, bf16 autocast in bf16 mode."""
        return self.enc.img_encoder.grid_features(input_img, dtype=eng.act_dtype)

    def classify_mask(self, attn_mask, eng):
        """[B,L,L] / [B,L] int mask -> (mode u8[B], t_len i32[B]) on the device, validated cell by cell."""
        import ctypes as C

        d = eng.dims
        m = attn_mask.to(device=eng.device, dtype=torch.int64).contiguous()
        B = m.shape[0]
        if m.dim() not in (2, 3) or m.shape[-1] != d.L:
            raise _lib.MedvillError("attn_mask must be [B, %d, %d] or [B, %d]" % (d.L, d.L, d.L))
        mode = torch.empty(B, dtype=torch.uint8, device=eng.device)
        t_len = torch.empty(B, dtype=torch.int32, device=eng.device)
        bad = torch.zeros(1, dtype=torch.int32, device=eng.device)
        _lib.check(_lib.lib().mv_mask_classify(_lib.ptr(m), m.dim(), B, d.A, d.L, _lib.ptr(mode), _lib.ptr(t_len), _lib.ptr(bad),
                                               _lib.stream_ptr(eng.device)), "mv_mask_classify")
        nbad = int(bad.item())
        if nbad:
            raise _lib.MedvillError("attention mask is not one of MedViLL's modes (Bidirectional / Seq2Seq / Bidirectional "
                                    "Auto-Regressive / Non-cross): %d cells differ from the predicate" % nbad)
        return mode, t_len

    def _encode(self, cls_tok, input_txt, attn_mask, segment, input_img, sep_tok, train, mode=None, t_len=None, txt_labels=None,
                is_aligned=None, feats=None, **kw):
        B = input_txt.shape[0]
        eng = self.engine(B)
        if feats is None:
            feats = self.grid_features(input_img.to(eng.device, non_blocking=True), eng)
        ridx = self.enc.img_encoder.sample_regions(feats.shape[1])
        if mode is None:
            mode, t_len = self.classify_mask(attn_mask, eng)
        self._dropout_step += 1
        seed = (int(getattr(self.args, "seed", 123)) << 32) ^ (self._dropout_step * 0x9E3779B1) ^ (eng.rank << 20)
        batch = eng.make_batch(cls_tok=cls_tok, input_ids=input_txt, segment=segment, sep_tok=sep_tok, mode=mode, t_len=t_len,
                               region_idx=ridx, feats=feats, txt_labels=txt_labels, is_aligned=is_aligned, seed=seed, train=train, **kw)
        eng.forward(batch)
        return eng, batch

    def forward(self, cls_tok, input_txt, attn_mask, segment, input_img, sep_tok):
        """-> (prediction_scores [B, L, V] fp32, itm_logits [B, 2] fp32).  Signature: cxrbert_origin.py:144.  In training mode
        with autograd enabled the outputs carry a grad_fn (see _JointForward); otherwise they are plain tensors."""
        if self._wants_grad():
            return _JointForward.apply(self._anchor(), self, "logits", cls_tok, input_txt, attn_mask, segment, input_img, sep_tok)
        B = int(input_txt.shape[0])
        cap = self.engine(min(B, int(getattr(self.args, "max_micro_batch", 64)))).max_batch
        outs = []
        for s0 in range(0, B, cap):                  # micro-batches of at most max_batch samples, as eval_step does
            sl = slice(s0, min(B, s0 + cap))
            eng, batch = self._encode(cls_tok[sl], input_txt[sl], attn_mask[sl], segment[sl], input_img[sl], sep_tok[sl], train=self.training)
            logits = eng.full_logits(batch)
            itm = torch.empty(batch.B, 2, dtype=torch.float32, device=eng.device)
            self._peek_into("itm_logits", 0, itm)
            outs.append((logits, itm))
        if len(outs) == 1:
            return outs[0]
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])

    def pretrain_step(self, cls_tok, input_ids, txt_labels, attn_masks, image, segment, is_aligned, sep_tok, lr=None,
                      mode=None, t_len=None, optimizer_step=True, feats=None, lazy=False):
        """One fused MLM+ITM step == models/train_origin.py:106-131 (forward, CE losses, zero_grad, backward, AdamW.step)
        plus the metrics of :133-146.  Tensors may live on the host (pinned) or the device.  Under torch.distributed
        (one process per GPU) gradients are all-reduced in buckets overlapped with backward and the loss normalisers
        are global, so N ranks x B samples reproduce a single-process batch of N*B.
        lazy=True returns a callable yielding the same dict (see below)."""
        B = int(input_ids.shape[0])
        eng = self.engine(min(B, int(getattr(self.args, "max_micro_batch", 64))))
        lab = torch.as_tensor(txt_labels)
        n_lab = int((lab != -100).sum())
        n_lab_g, b_g = float(n_lab), float(B)
        counts = None
        if eng.world > 1:
            # global loss normalisers stay on the device: all-reduced in-stream and read by the CE kernels through
            # mv_batch.global_counts, so the host never waits for the GPU here (a .tolist() drained the whole queue
            # once per step)
            counts = torch.tensor([n_lab_g, b_g], dtype=torch.float32).pin_memory().to(eng.device, non_blocking=True)
            eng.allreduce_f32(counts)
        if mode is None:
            mode, t_len = self.classify_mask(torch.as_tensor(attn_masks), eng)
        cap = eng.max_batch
        eng.stats_reset()
        chunks = [(s, min(B, s + cap)) for s in range(0, B, cap)]
        for ci, (s, e) in enumerate(chunks):
            sl = slice(s, e)
            _, batch = self._encode(cls_tok[sl], input_ids[sl], None, segment[sl], None if image is None else image[sl], sep_tok[sl],
                                    train=True, mode=mode[sl], t_len=t_len[sl], txt_labels=lab[sl], is_aligned=is_aligned[sl],
                                    feats=None if feats is None else feats[sl], n_lab_global=n_lab_g, batch_global=b_g,
                                    global_counts=counts)
            # gradients are exchanged only by the call that also steps the optimizer: with optimizer_step=False the local
            # sums stay local, so a later call adds to them and all-reduces ONCE (summed twice they would be scaled by world)
            eng.backward(batch, allreduce=(optimizer_step and eng.world > 1 and ci == len(chunks) - 1))
        if optimizer_step:
            eng.adamw_step(lr=float(self.args.lr if lr is None else lr))

        def finish(st):
            mlm = st["mlm_loss_sum"] / max(1, n_lab)
            itm = st["itm_loss_sum"] / B
            return dict(loss=mlm + itm, mlm_loss=mlm, itm_loss=itm, itm_correct=st["itm_correct"], mlm_correct=st["mlm_correct"],
                        n_labelled=n_lab, batch=B)
        if lazy:
            # the statistics travel to pinned host memory asynchronously; calling the returned object waits for that copy
            # only, so the caller can enqueue the next step first (the reference blocks on loss.item() every step)
            pending = eng.read_stats_async()
            return lambda: finish(pending())
        return finish(eng.read_stats())

    def eval_step(self, cls_tok, input_ids, txt_labels, attn_masks, image, segment, is_aligned, sep_tok, mode=None, t_len=None,
                  feats=None):
        """Validation step (models/train_origin.py:171-231): forward + both losses + accuracy counts, no gradients."""
        B = int(input_ids.shape[0])
        eng = self.engine(min(B, int(getattr(self.args, "max_micro_batch", 64))))
        lab = torch.as_tensor(txt_labels)
        n_lab = int((lab != -100).sum())
        if mode is None:
            mode, t_len = self.classify_mask(torch.as_tensor(attn_masks), eng)
        eng.stats_reset()
        for s in range(0, B, eng.max_batch):
            sl = slice(s, min(B, s + eng.max_batch))
            self._encode(cls_tok[sl], input_ids[sl], None, segment[sl], None if image is None else image[sl], sep_tok[sl],
                         train=False, mode=mode[sl], t_len=t_len[sl], txt_labels=lab[sl], is_aligned=is_aligned[sl],
                         feats=None if feats is None else feats[sl])
        st = eng.read_stats()
        mlm = st["mlm_loss_sum"] / max(1, n_lab)
        itm = st["itm_loss_sum"] / B
        return dict(loss=mlm + itm, mlm_loss=mlm, itm_loss=itm, itm_correct=st["itm_correct"], mlm_correct=st["mlm_correct"],
                    n_labelled=n_lab, batch=B)

    # -- HF-style persistence used by the trainer (models/train_origin.py:28-34,254-266) --
    def save_pretrained(self, save_directory):
        os.makedirs(save_directory, exist_ok=True)
        cfg = self.config.to_dict() if hasattr(self.config, "to_dict") else dict(self.config.__dict__)
        cfg = {k: v for k, v in cfg.items() if isinstance(v, (int, float, str, bool, type(None), list, dict))}
        with open(os.path.join(save_directory, "config.json"), "w") as f:
            json.dump(cfg, f, indent=2, sort_keys=True)
        torch.save({k: v.detach().cpu().clone() for k, v in self.state_dict().items()},
                   os.path.join(save_directory, "pytorch_model.bin"))

    OPTIMIZER_FILE = "optimizer.pt"

    def save_optimizer(self, save_directory):
        """Adam moments + step count + the dropout step counter next to pytorch_model.bin (absent in
        the reference, whose restart at train_origin.py:28-34 silently resets Adam)."""
        os.makedirs(save_directory, exist_ok=True)
        sd = self.engine().optimizer_state_dict()
        sd["dropout_step"] = int(self._dropout_step)
        torch.save(sd, os.path.join(save_directory, self.OPTIMIZER_FILE))

    def load_optimizer(self, directory):
        """Restore what save_optimizer wrote; call after .to(device).  Returns False when the checkpoint has no
        optimizer file (e.g. one written by the reference)."""
        path = os.path.join(directory, self.OPTIMIZER_FILE)
        if not os.path.isfile(path):
            return False
        sd = torch.load(path, map_location="cpu")
        self.engine().load_optimizer_state_dict(sd)
        self._dropout_step = int(sd.get("dropout_step", 0))
        return True

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, state_dict=None, config=None, args=None, **kw):
        if config is None:
            config = BertConfig.from_pretrained(pretrained_model_name_or_path)
        model = cls(config, args)
        if state_dict is None:
            state_dict = torch.load(os.path.join(pretrained_model_name_or_path, "pytorch_model.bin"), map_location="cpu")
        model.load_state_dict(state_dict, strict=False)
        return model
