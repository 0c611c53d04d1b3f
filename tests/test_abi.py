"""The C-ABI library loads and exports every symbol include/salt_b200.h declares; without a
GPU the compute entry points fail loudly instead of falling back to anything."""
import ctypes as C
import os
import re

import pytest

from salt_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return C.CDLL(api.LIB_PATH)


def test_exports_match_header(lib):
    hdr = open(os.path.join(ROOT, "include", "salt_b200.h")).read()
    names = sorted(set(re.findall(r"\b(salt_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert lib.salt_b200_abi_version() == 1


def test_host_layer_exports():
    """include/salt_host.h and the level-0 shim: every declared symbol is exported"""
    hdr = open(os.path.join(ROOT, "include", "salt_host.h")).read()
    names = sorted(set(re.findall(r"\b(salt_(?:chunk|pair|multi|fastq|sam|host)_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25 and "salt_sam_se" in names and "salt_fastq_split" in names and "salt_multi_init" in names
    H = C.CDLL(os.path.join(os.path.dirname(api.LIB_PATH), "libsalt_host.so"))
    for n in names:
        assert hasattr(H, n), n
    S = C.CDLL(os.path.join(os.path.dirname(api.LIB_PATH), "libsalt_level0.so"))
    for n in ("ed_mismatch", "ed_diff", "ed_diff_withcigar", "salt_level0_attach",        # editdistance.h:20-22
              "computeEditDistance", "computeEditDistanceWithCigar",                      # LandauVishkin.h:45, :50
              "ssw_init", "ssw_align", "init_destroy", "align_destroy"):                  # ssw.h:71, :111, :76, :124
        assert hasattr(S, n), n


def test_struct_sizes():
    assert api.PAIR_DT.itemsize == 8 and api.WIN_DT.itemsize == 12
    assert api.SSW_DT.itemsize == 28 and api.VERIFY_DT.itemsize == 16


def test_no_cpu_fallback(lib):
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present; the failure path is for CPU-only boxes")
    lib.salt_b200_init.restype = C.c_void_p
    lib.salt_b200_last_error.restype = C.c_char_p
    words = np.zeros(16, np.uint32)
    h = lib.salt_b200_init(words.ctypes.data_as(C.c_void_p), 128, None, 0, 0)
    assert not h
    assert b"no CUDA device" in lib.salt_b200_last_error()
    with pytest.raises(api.SaltError):
        api.Engine(words, 128)


def test_product_never_touches_oracle():
    """Nothing under salt_b200/ may import, link or call oracle/ or the emulator."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "salt_b200")):
        if "build" in dp.split(os.sep):
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                if re.search(r"(from|import)\s+oracle|liboracle|libsaltref|cuda_shim|emul_lib", txt):
                    bad.append(f)
    assert not bad, bad
