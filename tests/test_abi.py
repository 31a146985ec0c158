"""The C-ABI library loads and exports every symbol include/whippet_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "whippet_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(wq_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for s in ("wq_index_create", "wq_align_se", "wq_align_pe", "wq_counts_download", "wq_em_tpm", "wq_em_psi_batch"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from whippet_jl_b200.build import CUDA_SO
    assert os.path.exists(CUDA_SO)
    L = ctypes.CDLL(CUDA_SO)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert missing == []


def test_fails_loudly_without_gpu(fixture_index):
    import torch
    import whippet_jl_b200 as wq
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(wq.lib.WhippetB200Error):
        wq.GraphLibQuant(fixture_index)
