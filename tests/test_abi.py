"""The C-ABI library loads and exports every symbol include/bbgpu.h declares (no compute calls: CPU box)."""
import ctypes
import os
import re

import pytest

import barretenberg_b200 as bb
import helpers as H


def declared_symbols():
    text = open(os.path.join(H.ROOT, "include", "bbgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bbg_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("bbg_msm_g1", "bbg_msm_g1_batched", "bbg_ntt_fr", "bbg_ntt_fr_batched", "bbg_srs_register",
                 "bbg_generate_pippenger_point_table", "bbg_init", "bbg_plonk_create", "bbg_plonk_set_widgets",
                 "bbg_plonk_round_wires", "bbg_plonk_round_grand_product", "bbg_plonk_round_quotient", "bbg_plonk_round_evaluations",
                 "bbg_plonk_round_linearise", "bbg_plonk_round_openings", "bbg_fr_domain_lookup_table", "bbg_srs_from_transcript"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    path = bb.library_path()
    if not os.path.exists(path):
        pytest.skip("libbbgpu.so not built yet (python -c 'import __graft_entry__ as g; g.build()')")
    lib = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(lib, s), s


def test_no_gpu_means_loud_failure():
    """Without a CUDA device the product path must fail, not fall back."""
    path = bb.library_path()
    if not os.path.exists(path):
        pytest.skip("libbbgpu.so not built")
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(bb.BbgError):
        bb.Library(path)
