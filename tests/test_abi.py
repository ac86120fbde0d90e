"""The C-ABI library builds, loads and exports every symbol include/nph.h declares; without a GPU
nph_create fails loudly with NPH_ERR_NO_DEVICE (there is no CPU path)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nanopolish_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "nph.h")).read()
    declared = set(re.findall(r"\b(nph_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.nph_version() == 1


def test_pod_layouts_match_header():
    assert synth.READ_DT.itemsize == 64
    assert synth.HMM_JOB_DT.itemsize == 32
    assert synth.ABEA_JOB_DT.itemsize == 32
    assert synth.PAIR_DT.itemsize == 8
    assert synth.ABEA_RES_DT.itemsize == 24
    assert synth.HMM_JOB_DT.fields["stride"][1] == 28 and synth.HMM_JOB_DT.fields["flags"][1] == 30


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    lib = _lib.load()
    ctx = C.c_void_p()
    rc = lib.nph_create(C.byref(ctx), 0)
    assert rc == _lib.NPH_ERR_NO_DEVICE
    assert b"no CPU path" in lib.nph_strerror(rc)
    from nanopolish_b200.engine import Engine
    with pytest.raises(_lib.NphError):
        Engine(0)


def test_score_set_combine_is_host_arithmetic(port_oracle):
    lib = _lib.load()
    rng = np.random.default_rng(3)
    s = rng.uniform(-200, -100, 30).astype(np.float32)
    out = np.zeros(10, np.float32)
    assert lib.nph_score_set_combine(s.ctypes.data_as(C.c_void_p), 10, 3, out.ctypes.data_as(C.c_void_p)) == 0
    for g in range(10):
        want = port_oracle.score_set_combine(s[3 * g:3 * g + 3])
        assert np.float32(want).view(np.uint32) == out[g].view(np.uint32)
