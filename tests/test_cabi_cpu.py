"""CPU: the C-ABI library loads here (no GPU), exports every symbol include/medvill_sm100.h declares, answers the
host-only queries, and refuses to compute without an sm_100 device (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import medvill_b200 as m
from medvill_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "medvill_sm100.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "libmedvill_sm100.so lacks %s" % s
    assert set(syms) == set(_lib.SYMBOLS), set(syms) ^ set(_lib.SYMBOLS)      # ctypes table mirrors the header
    assert _lib.lib().mv_abi_version() == _lib.ABI_VERSION == 5


def test_layout_matches_reference_parameter_count():
    d = m.EngineDims()
    lay = m.query_layout(d)
    pm = m.param_map(d, lay)
    assert len(pm) == 208                                               # unique trainable tensors (aliases removed)
    assert sum(int(np.prod(s)) for _, s in pm.values()) == 111680060   # trainable parameter count
    spans = sorted((o, o + int(np.prod(s))) for o, s in pm.values())
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))          # no overlap
    assert spans[-1][1] <= lay["total"] and lay["total"] % 64 == 0
    # q|k|v are contiguous so one GEMM sees a [3H, H] weight and one [3H] bias
    H = d.hidden
    o_q, _ = pm["enc.encoder.layer.3.attention.self.query.weight"]
    o_k, _ = pm["enc.encoder.layer.3.attention.self.key.weight"]
    o_v, _ = pm["enc.encoder.layer.3.attention.self.value.weight"]
    assert (o_k - o_q, o_v - o_k) == (H * H, H * H)


def test_bucket_plan_partitions_the_arena_in_backward_order():
    d = m.EngineDims()
    lay = m.query_layout(d)
    b = m.bucket_plan(d)
    assert len(b) == d.layers + 2
    assert sum(c for _, c in b) == lay["total"]
    assert b[0][0] == lay["pool_w"] and b[-1] == (0, lay["layer0"])
    offs = [o for o, _ in b[1:-1]]
    assert offs == sorted(offs, reverse=True)                           # layer 11 first, layer 0 last
    covered = sorted((o, o + c) for o, c in b)
    assert covered[0][0] == 0 and all(x[1] == y[0] for x, y in zip(covered, covered[1:]))


def test_bad_dims_are_rejected_with_a_message():
    with pytest.raises(m.MedvillError, match="head dim must be 64"):
        m.query_layout(m.EngineDims(hidden=768, heads=8))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback():
    cfg = m.EngineDims().to_c(2, _lib.MV_PREC_BF16)
    h = C.c_void_p()
    rc = _lib.lib().mv_create(C.byref(h), C.byref(cfg))
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in _lib.lib().mv_last_error()
    with pytest.raises(m.MedvillError):
        m.PretrainEngine(m.EngineDims(), "cuda:0")


def test_ctypes_struct_layouts_match_the_c_header(tmp_path):
    """sizeof / offsetof of every struct in include/medvill_sm100.h, as gcc lays them out, equal the ctypes mirrors in
    _lib.py (a silent mismatch here would shift every pointer of an mv_batch)"""
    import ctypes as C
    import subprocess

    structs = {"mv_config": _lib.mv_config, "mv_layout": _lib.mv_layout, "mv_batch": _lib.mv_batch,
               "mv_step_stats": _lib.mv_step_stats, "mv_gemm_desc": _lib.mv_gemm_desc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "medvill_sm100.h"', 'int main(void) {']
    for name, cls in structs.items():
        lines.append('  printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for field, _ in cls._fields_:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, field, name, field))
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run(["gcc", "-I", inc, "-o", str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, cls in structs.items():
        assert int(got[name]) == C.sizeof(cls), name
        for field, _ in cls._fields_:
            assert int(got["%s.%s" % (name, field)]) == getattr(cls, field).offset, "%s.%s" % (name, field)
