"""TEST INFRASTRUCTURE — not product code.

Imports the *unmodified* reference modules from /root/reference under compatibility shims so that they
run on this container's transformers 5.x / torch 2.x, CPU only.  Used ONLY by oracle/make_golden.py (to pin
the oracle restatement against the real reference and to write tests/golden/*.npz).  /root/reference does not
exist on the GPU box, so nothing under tests/, bench.py or __graft_entry__.py imports this at run time.

Shims: old transformers module paths, offline BertConfig.from_pretrained,
BertEncoder tuple return, torchvision resnet50(pretrained=True) -> random init, Tensor.cuda no-op.
"""
import importlib.util
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("MEDVILL_REFERENCE", "/root/reference")

_state = {"config_kwargs": {}}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "cxrbert_origin.py"))


def set_bert_config(**kw):
    """kwargs that the shimmed BertConfig.from_pretrained() will pass to BertConfig()."""
    _state["config_kwargs"] = dict(kw)


def _install_shims():
    import transformers
    import transformers.models.bert.modeling_bert as mb
    from transformers import BertConfig

    if getattr(mb, "_medvill_shimmed", False):
        return
    # (1) transformers 3.x module paths used at models/cxrbert_origin.py:8-10
    legacy = types.ModuleType("transformers.modeling_bert")
    legacy.BertConfig = BertConfig
    legacy.BertModel = mb.BertModel
    legacy.BertPreTrainedModel = mb.BertPreTrainedModel
    sys.modules["transformers.modeling_bert"] = legacy
    auto = types.ModuleType("transformers.modeling_auto")
    auto.AutoModel = transformers.AutoModel
    auto.AutoConfig = transformers.AutoConfig
    sys.modules["transformers.modeling_auto"] = auto
    alb = types.ModuleType("transformers.modeling_albert")
    alb.AlbertModel = transformers.AlbertModel
    sys.modules["transformers.modeling_albert"] = alb
    tok_alb = types.ModuleType("transformers.tokenization_albert")
    tok_alb.AlbertTokenizer = object
    sys.modules["transformers.tokenization_albert"] = tok_alb

    # (2) offline config; eager attention (SDPA rejects the reference's fp16 additive mask)
    def _from_pretrained(cls, name=None, **kw):
        return cls(attn_implementation="eager", **_state["config_kwargs"])

    BertConfig.from_pretrained = classmethod(_from_pretrained)

    # (3) 3.x BertEncoder returned a tuple (last_hidden, attentions); cxrbert_origin.py:126-130 indexes [0],[1]
    _orig_forward = mb.BertEncoder.forward

    def _encoder_forward(self, hidden_states, attention_mask=None, output_hidden_states=False,
                         output_attentions=False, **kw):
        out = _orig_forward(self, hidden_states, attention_mask=attention_mask)
        last = out[0] if not torch.is_tensor(out) else out
        return (last, None)

    mb.BertEncoder.forward = _encoder_forward
    mb._medvill_shimmed = True

    # (4) no pretrained weights offline
    import torchvision

    _orig_resnet50 = torchvision.models.resnet50

    def _resnet50(pretrained=False, **kw):
        return _orig_resnet50(weights=None)

    torchvision.models.resnet50 = _resnet50

    # (5) hard-coded .cuda() calls (cxrbert_origin.py:92-95,115,117; image.py:60) on a CPU-only oracle
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self

    # dataset_origin.py:9 imports fuzzywuzzy (absent): equal iff token-sorted normalised strings match
    fw = types.ModuleType("fuzzywuzzy")
    fz = types.ModuleType("fuzzywuzzy.fuzz")

    def token_sort_ratio(a, b):
        import re

        def norm(s):
            return " ".join(sorted(re.sub(r"[^0-9a-z ]+", " ", str(s).lower()).split()))

        return 100 if norm(a) == norm(b) else 50

    fz.token_sort_ratio = token_sort_ratio
    fw.fuzz = fz
    sys.modules["fuzzywuzzy"] = fw
    sys.modules["fuzzywuzzy.fuzz"] = fz


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_models():
    """Returns (cxrbert_origin module, image module) imported from the reference tree, unchanged."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_shims()
    if "models" not in sys.modules or not hasattr(sys.modules["models"], "_medvill_ref"):
        pkg = types.ModuleType("models")
        pkg.__path__ = [os.path.join(REF_ROOT, "models")]
        pkg._medvill_ref = True
        sys.modules["models"] = pkg  # bypass models/__init__.py (pulls wandb + removed transformers.AdamW)
    image = _load("models.image", "models/image.py")
    cxr = _load("models.cxrbert_origin", "models/cxrbert_origin.py")
    return cxr, image


def load_reference_dataset():
    """Returns the reference data/dataset_origin.py module (CXRDataset), with the `self.disturbing_mask`
    AttributeError (dataset_origin.py:104) patched by a property reading args.disturbing_mask."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_shims()
    ds = _load("ref_dataset_origin", "data/dataset_origin.py")
    ds.CXRDataset.disturbing_mask = property(lambda s: s.args.disturbing_mask)
    return ds
