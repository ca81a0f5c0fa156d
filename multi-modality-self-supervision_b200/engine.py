"""Host-side wrapper of one `mv_handle`: owns the flat torch arenas (parameters, gradients, Adam moments, bf16 shadow)
that the C library borrows, maps the reference's state_dict names onto arena offsets, and marshals batches.

PyTorch is plumbing here (device memory, streams, torch.distributed rendezvous); every kernel of the step is in
libmedvill_sm100.so.  Reference surface mirrored: models/cxrbert_origin.py (parameter names), models/train_origin.py
(step order zero_grad -> backward -> step, losses, metrics).
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import MedvillError, check, lib, ptr, stream_ptr


@dataclass
class EngineDims:
    hidden: int = 768
    heads: int = 12
    layers: int = 12
    inter: int = 3072
    vocab: int = 30522
    max_pos: int = 512
    type_vocab: int = 2
    num_image_embeds: int = 180
    seq_len: int = 253
    img_hidden: int = 2048
    grid: int = 256
    ln_eps: float = 1e-12
    head_ln_eps: float = 1e-5
    dropout_p: float = 0.1          # hidden-state dropout (BertConfig.hidden_dropout_prob)
    attn_dropout_p: float = None    # attention-probability dropout (None: same as dropout_p)
    img_dropout_p: float = None     # image-embedding dropout, args.dropout_prob (None: same as dropout_p)
    flags: int = 0                  # _lib.FLAG_*

    @property
    def A(self):
        return self.num_image_embeds + 2

    @property
    def T(self):
        return self.seq_len + 1

    @property
    def L(self):
        return self.num_image_embeds + self.seq_len + 3

    def to_c(self, max_batch, precision):
        return _lib.mv_config(hidden=self.hidden, heads=self.heads, layers=self.layers, inter=self.inter, vocab=self.vocab,
                              max_pos=self.max_pos, type_vocab=self.type_vocab, num_image_embeds=self.num_image_embeds,
                              seq_len=self.seq_len, img_hidden=self.img_hidden, grid=self.grid, max_batch=max_batch,
                              precision=precision, ln_eps=self.ln_eps, head_ln_eps=self.head_ln_eps, dropout_p=self.dropout_p,
                              attn_dropout_p=self.dropout_p if self.attn_dropout_p is None else self.attn_dropout_p,
                              img_dropout_p=self.dropout_p if self.img_dropout_p is None else self.img_dropout_p, flags=int(self.flags))


def query_layout(dims, max_batch=1, precision=_lib.MV_PREC_BF16):
    """Arena layout for `dims` (pure host call: works without a GPU)."""
    cfg = dims.to_c(max_batch, precision)
    lay = _lib.mv_layout()
    check(lib().mv_layout_query(C.byref(cfg), C.byref(lay)), "mv_layout_query")
    return {n: getattr(lay, n) for n in _lib.LAYOUT_FIELDS}


def bucket_plan(dims):
    """Gradient all-reduce buckets [(offset, count)] in the order backward completes them (host call)."""
    cfg = dims.to_c(1, _lib.MV_PREC_BF16)
    offs = (C.c_int64 * 64)()
    cnts = (C.c_int64 * 64)()
    n = C.c_int32(0)
    check(lib().mv_bucket_plan(C.byref(cfg), offs, cnts, 64, C.byref(n)), "mv_bucket_plan")
    return [(offs[i], cnts[i]) for i in range(n.value)]


def param_map(dims, lay):
    """reference state_dict key -> (arena offset, shape) for the 111.68 M trainable parameters."""
    H, I, V = dims.hidden, dims.inter, dims.vocab
    m = {
        "enc.txt_embeddings.word_embeddings.weight": (lay["word"], (V, H)),
        "enc.txt_embeddings.position_embeddings.weight": (lay["pos"], (dims.max_pos, H)),
        "enc.txt_embeddings.token_type_embeddings.weight": (lay["type"], (dims.type_vocab, H)),
        "enc.txt_embeddings.LayerNorm.weight": (lay["emb_ln_g"], (H,)),
        "enc.txt_embeddings.LayerNorm.bias": (lay["emb_ln_b"], (H,)),
        "enc.img_embeddings.img_embeddings.weight": (lay["img_w"], (H, dims.img_hidden)),
        "enc.img_embeddings.img_embeddings.bias": (lay["img_b"], (H,)),
        "enc.pooler.dense.weight": (lay["pool_w"], (H, H)),
        "enc.pooler.dense.bias": (lay["pool_b"], (H,)),
        "mlm.predictions.bias": (lay["mlm_bias"], (V,)),
        "mlm.predictions.transform.dense.weight": (lay["mlm_tw"], (H, H)),
        "mlm.predictions.transform.dense.bias": (lay["mlm_tb"], (H,)),
        "mlm.predictions.transform.LayerNorm.weight": (lay["mlm_ln_g"], (H,)),
        "mlm.predictions.transform.LayerNorm.bias": (lay["mlm_ln_b"], (H,)),
        "itm.linear.weight": (lay["itm_w"], (2, H)),
        "itm.linear.bias": (lay["itm_b"], (2,)),
    }
    for l in range(dims.layers):
        base = lay["layer0"] + l * lay["layer_stride"]
        p = "enc.encoder.layer.%d." % l
        for j, n in enumerate(("query", "key", "value")):
            m[p + "attention.self.%s.weight" % n] = (base + lay["l_wqkv"] + j * H * H, (H, H))
            m[p + "attention.self.%s.bias" % n] = (base + lay["l_bqkv"] + j * H, (H,))
        m[p + "attention.output.dense.weight"] = (base + lay["l_wo"], (H, H))
        m[p + "attention.output.dense.bias"] = (base + lay["l_bo"], (H,))
        m[p + "attention.output.LayerNorm.weight"] = (base + lay["l_ln1_g"], (H,))
        m[p + "attention.output.LayerNorm.bias"] = (base + lay["l_ln1_b"], (H,))
        m[p + "intermediate.dense.weight"] = (base + lay["l_w1"], (I, H))
        m[p + "intermediate.dense.bias"] = (base + lay["l_b1"], (I,))
        m[p + "output.dense.weight"] = (base + lay["l_w2"], (H, I))
        m[p + "output.dense.bias"] = (base + lay["l_b2"], (H,))
        m[p + "output.LayerNorm.weight"] = (base + lay["l_ln2_g"], (H,))
        m[p + "output.LayerNorm.bias"] = (base + lay["l_ln2_b"], (H,))
    return m


class Batch:
    """One micro-batch resident on the device, plus the mv_batch struct pointing into it."""

    def __init__(self, engine, cls_tok, input_ids, segment, sep_tok, mode, t_len, region_idx, feats, txt_labels=None,
                 is_aligned=None, lab_rows=None, lab_labels=None, n_lab_global=None, batch_global=None, seed=0, train=True,
                 sep_position=0, prefix_type=0, pad_lookup_grad=False, lab_weights=None, global_counts=None, drop_worst_keep=0):
        d = engine.dims
        dev = engine.device
        i64 = lambda t: None if t is None else torch.as_tensor(t).to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
        self.cls_tok, self.sep_tok = i64(cls_tok).reshape(-1), i64(sep_tok).reshape(-1)
        self.input_ids, self.segment = i64(input_ids), i64(segment)
        self.B = int(self.input_ids.shape[0])
        if tuple(self.input_ids.shape) != (self.B, d.T) or tuple(self.segment.shape) != (self.B, d.T):
            raise MedvillError("input_ids/segment must be [B, %d], got %s" % (d.T, tuple(self.input_ids.shape)))
        self.is_aligned = i64(is_aligned)
        self.region_idx = i64(region_idx)
        if self.region_idx.numel() != d.num_image_embeds:
            raise MedvillError("region_idx must hold %d grid positions" % d.num_image_embeds)
        self.mode = torch.as_tensor(mode).to(device=dev, dtype=torch.uint8, non_blocking=True).contiguous()
        self.t_len = torch.as_tensor(t_len).to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
        self.feats = feats.to(device=dev, dtype=engine.act_dtype).contiguous()
        if tuple(self.feats.shape) != (self.B, d.grid, d.img_hidden):
            raise MedvillError("feats must be [B, %d, %d], got %s" % (d.grid, d.img_hidden, tuple(self.feats.shape)))
        if lab_rows is None and txt_labels is not None:
            # positions with a label (txt_labels != -100), flattened b * L + s  (train_origin.py:62 ignore_index=-100)
            lab = torch.as_tensor(txt_labels)
            if lab.is_cuda:
                flat = lab.reshape(-1)
                lab_rows = torch.nonzero(flat != -100).reshape(-1)
                lab_labels = flat[lab_rows]
            else:
                flat = lab.reshape(-1).numpy()
                rows = np.flatnonzero(flat != -100)
                lab_rows, lab_labels = torch.from_numpy(rows), torch.from_numpy(flat[rows])
        self.lab_rows, self.lab_labels = i64(lab_rows), i64(lab_labels)
        self.n_lab = 0 if self.lab_rows is None else int(self.lab_rows.numel())
        self.global_counts = global_counts          # device float32 [2] (kept alive here) or None
        self.lab_weights = None if lab_weights is None else torch.as_tensor(lab_weights).to(device=dev, dtype=torch.float32).contiguous()
        if self.lab_weights is not None and self.lab_weights.numel() != self.n_lab:
            raise MedvillError("lab_weights must hold one weight per labelled row")
        n_glob = self.n_lab if n_lab_global is None else n_lab_global
        b_glob = self.B if batch_global is None else batch_global
        self.c = _lib.mv_batch(
            B=self.B, cls_tok=ptr(self.cls_tok), sep_tok=ptr(self.sep_tok), input_ids=ptr(self.input_ids),
            segment=ptr(self.segment), is_aligned=ptr(self.is_aligned), region_idx=ptr(self.region_idx),
            mode=ptr(self.mode), t_len=ptr(self.t_len), feats=ptr(self.feats), n_lab=self.n_lab,
            lab_rows=ptr(self.lab_rows) if self.n_lab else None, lab_labels=ptr(self.lab_labels) if self.n_lab else None,
            inv_n_lab_global=1.0 / max(1, n_glob), inv_batch_global=1.0 / max(1, b_glob), dropout_seed=int(seed) & (2 ** 64 - 1),
            train=1 if train else 0, sep_position=int(sep_position), prefix_type=int(prefix_type),
            pad_lookup_grad=1 if pad_lookup_grad else 0, global_counts=ptr(self.global_counts), lab_weights=ptr(self.lab_weights),
            drop_worst_keep=int(drop_worst_keep))


def _raise_on_error_flags(flags):
    """mv_step_stats.error_flags: the kernels clamp an out-of-range id and report it; PyTorch raises IndexError there."""
    if flags:
        what = [n for bit, n in ((_lib.ERR_TOKEN_ID, "token id >= vocab_size"), (_lib.ERR_SEGMENT_ID, "segment id >= type_vocab_size"),
                                 (_lib.ERR_REGION_IDX, "region index >= max_position_embeddings"),
                                 (_lib.ERR_MLM_LABEL, "MLM label outside [0, vocab_size)")) if flags & bit]
        raise MedvillError("index out of range in the last step: " + "; ".join(what))


_LIVE_ENGINES = []      # weak references; lets optimizers locate the engine that owns a parameter view


def find_engine(tensor):
    """The live engine whose parameter arena contains `tensor` (an nn.Parameter view), or None."""
    ptr_ = tensor.data_ptr()
    for ref in list(_LIVE_ENGINES):
        eng = ref()
        if eng is None or not getattr(eng, "_h", None) or not eng._h.value:
            _LIVE_ENGINES.remove(ref)
            continue
        lo = eng.params.data_ptr()
        if lo <= ptr_ < lo + eng.params.numel() * 4:
            return eng
    return None


class PretrainEngine:
    """The B200 engine for one rank: arenas + handle + step drivers."""

    def __init__(self, dims, device, precision="bf16", max_batch=64):
        if not torch.cuda.is_available():
            raise MedvillError("CUDA device required: the MedViLL sm_100a engine has no CPU fallback")
        self.dims = dims
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise MedvillError("PretrainEngine needs a cuda device, got %s" % device)
        self.precision = _lib.MV_PREC_FP32 if precision in ("fp32", "float32", _lib.MV_PREC_FP32) else _lib.MV_PREC_BF16
        self.act_dtype = torch.float32 if self.precision == _lib.MV_PREC_FP32 else torch.bfloat16
        self.max_batch = max_batch
        self.layout = query_layout(dims, max_batch, self.precision)
        self.pmap = param_map(dims, self.layout)
        n = self.layout["total"]
        torch.cuda.set_device(self.device)
        z = lambda dt: torch.zeros(n, dtype=dt, device=self.device)
        self.params, self.grads, self.adam_m, self.adam_v = z(torch.float32), z(torch.float32), z(torch.float32), z(torch.float32)
        self.shadow = z(torch.bfloat16) if self.precision == _lib.MV_PREC_BF16 else None
        self._h = C.c_void_p()
        cfg = dims.to_c(max_batch, self.precision)
        check(lib().mv_create(C.byref(self._h), C.byref(cfg)), "mv_create")
        check(lib().mv_bind_arenas(self._h, ptr(self.params), ptr(self.grads), ptr(self.adam_m), ptr(self.adam_v), ptr(self.shadow)),
              "mv_bind_arenas")
        self.step_count = 0
        self.world, self.rank = 1, 0
        import weakref

        _LIVE_ENGINES.append(weakref.ref(self))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().mv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters ----
    def view(self, name, arena=None):
        off, shape = self.pmap[name]
        a = self.params if arena is None else arena
        n = int(np.prod(shape))
        return a[off:off + n].view(shape)

    def load_params(self, named):
        """Copy reference-named fp32 tensors into the arena and refresh the bf16 shadow."""
        with torch.no_grad():
            for name in self.pmap:
                if name in named:
                    self.view(name).copy_(torch.as_tensor(named[name]).to(self.device, torch.float32))
        self.refresh_shadow()

    def refresh_shadow(self):
        check(lib().mv_refresh_shadow(self._h, stream_ptr(self.device)), "mv_refresh_shadow")

    # ---- the step ----
    def make_batch(self, **kw):
        return Batch(self, **kw)

    def stats_reset(self):
        check(lib().mv_stats_reset(self._h, stream_ptr(self.device)), "mv_stats_reset")

    def forward(self, batch):
        check(lib().mv_forward(self._h, C.byref(batch.c), stream_ptr(self.device)), "mv_forward")

    def backward(self, batch, allreduce=False):
        check(lib().mv_backward(self._h, C.byref(batch.c), 1 if allreduce else 0, stream_ptr(self.device)), "mv_backward")

    def backward_external(self, batch, rows=None, dlogits=None, d_itm=None, d_seq=None, d_pooled=None, allreduce=False):
        """Backward from output gradients computed by the caller (torch.autograd drop-in path); tensors on the device."""
        n = 0 if rows is None else int(rows.numel())
        g = _lib.mv_external_grads(n_rows=n, rows=ptr(rows) if n else None, dlogits=ptr(dlogits) if n else None, d_itm=ptr(d_itm),
                                   d_seq=ptr(d_seq), d_pooled=ptr(d_pooled))
        check(lib().mv_backward_external(self._h, C.byref(batch.c), C.byref(g), 1 if allreduce else 0, stream_ptr(self.device)),
              "mv_backward_external")

    def zero_grads(self):
        check(lib().mv_zero_grads(self._h, stream_ptr(self.device)), "mv_zero_grads")

    def adamw_step(self, lr, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, grad_scale=1.0):
        """HF-3.x AdamW defaults as instantiated at models/train_origin.py:60 (only lr is passed there)."""
        self.step_count += 1
        check(lib().mv_adamw_step(self._h, lr, betas[0], betas[1], eps, weight_decay, self.step_count, grad_scale,
                                  stream_ptr(self.device)), "mv_adamw_step")

    def optimizer_state_dict(self):
        """Adam moments keyed by reference parameter name (torch.optim layout: step / exp_avg / exp_avg_sq), host tensors.
        The reference saves no optimizer state (models/train_origin.py:254-266), so a restart there resets Adam."""
        torch.cuda.synchronize(self.device)
        return {"step": int(self.step_count),
                "state": {n: {"exp_avg": self.view(n, self.adam_m).detach().cpu().clone(),
                              "exp_avg_sq": self.view(n, self.adam_v).detach().cpu().clone()} for n in self.pmap}}

    def load_optimizer_state_dict(self, sd):
        missing = sorted(set(self.pmap) - set(sd["state"]))
        if missing:
            raise MedvillError("optimizer state lacks %d tensors, e.g. %s" % (len(missing), missing[:3]))
        with torch.no_grad():
            for n in self.pmap:
                for key, arena in (("exp_avg", self.adam_m), ("exp_avg_sq", self.adam_v)):
                    src, dst = sd["state"][n][key], self.view(n, arena)
                    if tuple(src.shape) != tuple(dst.shape):
                        raise MedvillError("optimizer state %s.%s has shape %s, expected %s" % (n, key, tuple(src.shape), tuple(dst.shape)))
                    dst.copy_(src.to(device=self.device, dtype=torch.float32))
        self.step_count = int(sd["step"])

    def bert_adam_step(self, lr, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01, max_grad_norm=1.0):
        """BertAdam.step (fine-tune; .../sc/pytorch_pretrained_bert/optimization.py:112-182); `lr` is the scheduled rate."""
        self.step_count += 1
        check(lib().mv_bert_adam_step(self._h, lr, betas[0], betas[1], eps, weight_decay, max_grad_norm, stream_ptr(self.device)),
              "mv_bert_adam_step")

    def read_stats(self):
        st = _lib.mv_step_stats()
        check(lib().mv_read_stats(self._h, C.byref(st), stream_ptr(self.device)), "mv_read_stats")
        _raise_on_error_flags(st.error_flags)
        return dict(mlm_loss_sum=st.mlm_loss_sum, itm_loss_sum=st.itm_loss_sum, mlm_correct=st.mlm_correct, itm_correct=st.itm_correct)

    def read_stats_async(self):
        """Enqueue the statistics read-back into a rotating pinned slot; returns a zero-argument callable that waits for
        that copy (only) and yields the same dict as read_stats().  Lets a trainer keep the GPU queue full and look at
        the loss one step (or `log_freq` steps) later."""
        if not hasattr(self, "_stat_slots"):
            self._stat_slots = [torch.zeros(8, dtype=torch.int32).pin_memory() for _ in range(8)]
            self._stat_next = 0
        slot = self._stat_slots[self._stat_next % len(self._stat_slots)]
        self._stat_next += 1
        check(lib().mv_read_stats_async(self._h, ptr(slot), stream_ptr(self.device)), "mv_read_stats_async")
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))

        def resolve():
            ev.synchronize()
            f = slot.view(torch.float32)
            _raise_on_error_flags(int(slot[4]))
            return dict(mlm_loss_sum=float(f[0]), itm_loss_sum=float(f[1]), mlm_correct=int(slot[2]), itm_correct=int(slot[3]))
        return resolve

    def itm_logits(self, B):
        out = torch.empty(B, 2, dtype=torch.float32)
        check(lib().mv_itm_logits(self._h, ptr(out), B, stream_ptr(self.device)), "mv_itm_logits")
        return out

    def full_logits(self, batch):
        """[B, L, V] fp32 prediction scores for every position (reference CXRBERT.forward output)."""
        d = self.dims
        ld = self.layout["vocab_padded"]
        out = torch.empty(batch.B * d.L, ld, dtype=torch.float32, device=self.device)
        check(lib().mv_full_logits(self._h, C.byref(batch.c), ptr(out), ld, stream_ptr(self.device)), "mv_full_logits")
        return out.view(batch.B, d.L, ld)[:, :, :d.vocab]

    def peek(self, name, layer=0, shape=None, dtype=None):
        """Copy a named intermediate (parity-test aid); returns a CPU tensor."""
        dtype = dtype or self.act_dtype
        n = int(np.prod(shape))
        out = torch.empty(n, dtype=dtype)
        got = C.c_int64(0)
        check(lib().mv_peek(self._h, name.encode(), layer, ptr(out), out.numel() * out.element_size(), C.byref(got),
                            stream_ptr(self.device)), "mv_peek")
        return out.view(shape)

    # ---- data parallel ----
    def comm_init(self, rank, world, store_broadcast):
        """`store_broadcast(bytes_or_None) -> bytes` distributes rank 0's 128-byte NCCL unique id (torch.distributed plumbing)."""
        uid = (C.c_uint8 * 128)()
        if rank == 0:
            check(lib().mv_comm_unique_id(uid), "mv_comm_unique_id")
        raw = store_broadcast(bytes(uid) if rank == 0 else None)
        uid = (C.c_uint8 * 128).from_buffer_copy(raw)
        check(lib().mv_comm_init(self._h, uid, rank, world), "mv_comm_init")
        self.rank, self.world = rank, world

    def comm_sync(self):
        """Make the current stream wait for every pending gradient-bucket all-reduce."""
        check(lib().mv_comm_sync(self._h, stream_ptr(self.device)), "mv_comm_sync")

    def allreduce_f32(self, t):
        check(lib().mv_comm_allreduce_f32(self._h, ptr(t), t.numel(), stream_ptr(self.device)), "mv_comm_allreduce_f32")
        return t
