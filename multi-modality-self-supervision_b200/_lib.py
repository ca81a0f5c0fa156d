"""ctypes binding of libmedvill_sm100.so (declarations mirror include/medvill_sm100.h one to one).

The library is the product: there is no Python / PyTorch / CPU fallback for any compute entry.  If the shared
object is missing or a call fails, `MedvillError` is raised with the library's own message.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmedvill_sm100.so")
CSRC = os.path.join(_HERE, "csrc")

ABI_VERSION = 5          # must equal MV_ABI_VERSION of include/medvill_sm100.h (struct layouts below mirror that header)
MV_PREC_BF16, MV_PREC_FP32 = 0, 1
MODE_BIDIR, MODE_S2S, MODE_BAR, MODE_NONCROSS, MODE_S2S_FT, MODE_BAR_FT = 0, 1, 2, 3, 4, 5
EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESID, EPI_BIAS_TANH, EPI_RESID, EPI_DGELU, EPI_BIAS_GELU_GRAD, EPI_MUL = range(9)


class MedvillError(RuntimeError):
    pass


class mv_config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("hidden", "heads", "layers", "inter", "vocab", "max_pos", "type_vocab",
                                         "num_image_embeds", "seq_len", "img_hidden", "grid", "max_batch", "precision")] + \
               [("ln_eps", C.c_float), ("head_ln_eps", C.c_float), ("dropout_p", C.c_float), ("attn_dropout_p", C.c_float),
                ("img_dropout_p", C.c_float), ("flags", C.c_int32)]


FLAG_BF16_RESIDUAL, FLAG_DETERMINISTIC = 1, 2
ERR_TOKEN_ID, ERR_SEGMENT_ID, ERR_REGION_IDX, ERR_MLM_LABEL = 1, 2, 4, 8


LAYOUT_FIELDS = ("total", "word", "pos", "type", "emb_ln_g", "emb_ln_b", "img_w", "img_b", "layer0", "layer_stride",
                 "l_wqkv", "l_bqkv", "l_wo", "l_bo", "l_ln1_g", "l_ln1_b", "l_w1", "l_b1", "l_w2", "l_b2", "l_ln2_g", "l_ln2_b",
                 "pool_w", "pool_b", "mlm_bias", "mlm_tw", "mlm_tb", "mlm_ln_g", "mlm_ln_b", "itm_w", "itm_b", "vocab_padded")


class mv_layout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in LAYOUT_FIELDS]


class mv_batch(C.Structure):
    _fields_ = [("B", C.c_int32), ("cls_tok", C.c_void_p), ("sep_tok", C.c_void_p), ("input_ids", C.c_void_p),
                ("segment", C.c_void_p), ("is_aligned", C.c_void_p), ("region_idx", C.c_void_p), ("mode", C.c_void_p),
                ("t_len", C.c_void_p), ("feats", C.c_void_p), ("n_lab", C.c_int32), ("lab_rows", C.c_void_p),
                ("lab_labels", C.c_void_p), ("inv_n_lab_global", C.c_float), ("inv_batch_global", C.c_float),
                ("dropout_seed", C.c_uint64), ("train", C.c_int32),
                ("sep_position", C.c_int32), ("prefix_type", C.c_int32), ("pad_lookup_grad", C.c_int32),
                ("global_counts", C.c_void_p), ("lab_weights", C.c_void_p), ("drop_worst_keep", C.c_int32)]


class mv_step_stats(C.Structure):
    _fields_ = [("mlm_loss_sum", C.c_float), ("itm_loss_sum", C.c_float), ("mlm_correct", C.c_int32), ("itm_correct", C.c_int32),
                ("error_flags", C.c_int32)]


class mv_external_grads(C.Structure):
    _fields_ = [("n_rows", C.c_int32), ("rows", C.c_void_p), ("dlogits", C.c_void_p), ("d_itm", C.c_void_p), ("d_seq", C.c_void_p),
                ("d_pooled", C.c_void_p)]


class mv_gemm_desc(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
                ("A", C.c_void_p), ("lda", C.c_int64), ("a_mn", C.c_int32),
                ("B", C.c_void_p), ("ldb", C.c_int64), ("b_mn", C.c_int32),
                ("C", C.c_void_p), ("ldc", C.c_int64), ("c_f32", C.c_int32), ("accumulate", C.c_int32),
                ("C2", C.c_void_p), ("ldc2", C.c_int64), ("epi", C.c_int32), ("bias", C.c_void_p),
                ("resid", C.c_void_p), ("ldr", C.c_int64), ("aux", C.c_void_p), ("ldaux", C.c_int64),
                ("dropout_p", C.c_float), ("dropout_seed", C.c_uint64), ("dropout_site", C.c_uint32), ("resid_f32", C.c_int32)]


# every symbol include/medvill_sm100.h declares: (restype, argtypes)
_P, _I, _L, _F, _U64, _U32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64, C.c_uint32
SYMBOLS = {
    "mv_last_error": (C.c_char_p, []),
    "mv_abi_version": (_I, []),
    "mv_layout_query": (_I, [C.POINTER(mv_config), C.POINTER(mv_layout)]),
    "mv_bucket_plan": (_I, [C.POINTER(mv_config), C.POINTER(_L), C.POINTER(_L), _I, C.POINTER(_I)]),
    "mv_create": (_I, [C.POINTER(_P), C.POINTER(mv_config)]),
    "mv_destroy": (_I, [_P]),
    "mv_bind_arenas": (_I, [_P, _P, _P, _P, _P, _P]),
    "mv_refresh_shadow": (_I, [_P, _P]),
    "mv_stats_reset": (_I, [_P, _P]),
    "mv_forward": (_I, [_P, C.POINTER(mv_batch), _P]),
    "mv_backward": (_I, [_P, C.POINTER(mv_batch), _I, _P]),
    "mv_backward_external": (_I, [_P, C.POINTER(mv_batch), C.POINTER(mv_external_grads), _I, _P]),
    "mv_zero_grads": (_I, [_P, _P]),
    "mv_adamw_step": (_I, [_P, _F, _F, _F, _F, _F, _I, _F, _P]),
    "mv_bert_adam_step": (_I, [_P, _F, _F, _F, _F, _F, _F, _P]),
    "mv_read_stats": (_I, [_P, C.POINTER(mv_step_stats), _P]),
    "mv_read_stats_async": (_I, [_P, _P, _P]),
    "mv_itm_logits": (_I, [_P, _P, _I, _P]),
    "mv_itm_match_prob": (_I, [_P, _P, _I, _P]),
    "mv_full_logits": (_I, [_P, C.POINTER(mv_batch), _P, _L, _P]),
    "mv_peek": (_I, [_P, C.c_char_p, _I, _P, _L, C.POINTER(_L), _P]),
    "mv_launch_count": (C.c_long, []),
    "mv_profile": (_I, [_P, _I]),
    "mv_profile_read": (_I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_I), _I]),
    "mv_comm_unique_id": (_I, [_P]),
    "mv_comm_init": (_I, [_P, _P, _I, _I]),
    "mv_comm_allreduce_f32": (_I, [_P, _P, _L, _P]),
    "mv_comm_sync": (_I, [_P, _P]),
    "mv_gemm": (_I, [C.POINTER(mv_gemm_desc), _I, _P]),
    "mv_attn_mask_dump": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "mv_mask_classify": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "mv_attention_fwd": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P, _F, _U64, _U32, _I, _P]),
    "mv_attention_bwd": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _U64, _U32, _I, _P]),
    "mv_layernorm_fwd": (_I, [_P, _P, _P, _P, _I, _I, _F, _I, _P]),
    "mv_layernorm_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _F, _I, _P]),
    "mv_mlm_ce": (_I, [_P, _L, _P, _I, _I, _P, _F, _P, _P, _P, _P, _I, _P]),
    "mv_bn_workspace_floats": (_L, [_L, _I]),
    "mv_bn_forward": (_I, [_P, _P, _P, _L, _I, _P, _P, _P, _P, _F, _F, _I, _I, _P, _L, _I, _P]),
    "mv_normalize_u8": (_I, [_P, _P, _L, _L, _I, C.POINTER(C.c_float), C.POINTER(C.c_float), _I, _P]),
    "mv_normalize_u8_s2d": (_I, [_P, _P, _I, _I, _I, C.POINTER(C.c_float), C.POINTER(C.c_float), _I, _P]),
    "mv_bn_relu_maxpool": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _F, _F, _I, _P, _L, _I, _P]),
    "mv_itm_head": (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _I, _P]),
    "mv_stem_conv_s2d": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "mv_adamw": (_I, [_P, _P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _F, _I, _P]),
}

_lib = None


def build(verbose=False):
    """Compile the CUDA sources for sm_100a (nvcc cross-compiles without a GPU) into libmedvill_sm100.so."""
    out = subprocess.run(["make", "-C", CSRC, "-j", str(os.cpu_count() or 4), "all", "tests"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise MedvillError("building libmedvill_sm100.so failed (see output above)")
    return LIB_PATH


def lib():
    """The loaded library; raises MedvillError (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise MedvillError("%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(or make -C %s). There is no fallback implementation." % (LIB_PATH, CSRC))
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        if l.mv_abi_version() != ABI_VERSION:
            raise MedvillError("libmedvill_sm100.so ABI version mismatch")
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().mv_last_error()
        raise MedvillError("%s failed (status %d): %s" % (what or "libmedvill_sm100 call", rc, msg.decode() if msg else "?"))


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    import torch

    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
