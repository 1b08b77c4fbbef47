"""CPU: the C-ABI library loads, exports every symbol include/deepof_b200.h declares, and its
host-side logic (state layout, graph operators) matches the reference golden files.
No compute entry point is called (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import golden_cases, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    from deepof_b200 import _lib
    return _lib


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "deepof_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dof_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    raw = C.CDLL(built.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    assert declared == set(built.EXPORTS), declared ^ set(built.EXPORTS)
    assert built.lib().dof_abi_version() == 3


@pytest.mark.parametrize("case", golden_cases())
def test_state_layout_is_reference_state_dict(built, case):
    from deepof_b200.vade import state_layout
    g = load_golden(case)
    d = g["dims"]
    cfg = built.DofConfig(d["T"], d["N"], d["E"], 3, 1, d["D"], d["K"])
    lay = state_layout(cfg)
    names = [k[2:] for k in g if k.startswith("p/")]
    assert [l[0] for l in lay] == names
    off = 0
    for name, o, numel, shape, grp in lay:
        assert tuple(g["p/" + name].shape) == shape, name
        assert o == off
        off += numel
        has_grad = ("g/" + name) in g
        assert (grp > 0) == has_grad, (name, grp)   # dead parameters / buffers are group 0
    assert off == built.lib().dof_state_numel(C.byref(cfg))


@pytest.mark.parametrize("case", golden_cases())
def test_graph_operators_match_reference(built, case):
    from deepof_b200.vade import graph_operators
    g = load_golden(case)
    lap, elap, inc = graph_operators(g["adjacency"])
    assert np.array_equal(inc, g["p/encoder.incidence"])
    np.testing.assert_allclose(lap, g["p/encoder.laplacian"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(elap, g["p/encoder.edge_laplacian"], rtol=0, atol=1e-7)


def test_bad_config_is_rejected(built):
    cfg = built.DofConfig(25, 14, 14, 3, 1, 16, 64)   # K > 32
    assert built.lib().dof_state_numel(C.byref(cfg)) < 0
    assert b"n_components" in built.lib().dof_last_error()
    assert built.lib().dof_workspace_bytes(C.byref(cfg), 16, 1) == 0


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "deepof_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"import\s+oracle|from\s+oracle|oracle[/.]\w", src), f


def test_transformer_layouts_are_reference_state_dict(built):
    """dof_tfm_entry / dof_tfm_dec_entry (host-only) enumerate the float tensors of TFMEncoderPT / TFMDecoderPT in the
    reference's state_dict order with its shapes (goldens tfm_*.npz, tfmmodel_*.npz)."""
    from helpers import golden_cases_of, load_golden_of
    from deepof_b200._lib import DofTfmCfg, DofTfmDecCfg
    from deepof_b200.tfm import tfm_layout
    L = built.lib()
    for case in golden_cases_of("tfm"):
        g = load_golden_of("tfm", case)
        T, N, E, D, B, dk, heads, dff, layers = (int(v) for v in g["meta"])
        cfg = DofTfmCfg(T, N, E, 3, 1, D, dk, heads, dff, layers)
        lay = tfm_layout(cfg)
        names = [k[2:] for k in g if k.startswith("p/") and "num_batches_tracked" not in k]
        assert [l[0] for l in lay] == names
        off = 0
        for name, o, numel, shape in lay:
            assert tuple(g["p/" + name].shape) == shape and o == off, name
            off += numel
        assert off == L.dof_tfm_numel(C.byref(cfg))
        assert L.dof_tfm_workspace_bytes(C.byref(cfg), 64) > 0
    g = load_golden_of("tfmmodel", "vade")
    T, N, E, D, K, B = (int(v) for v in g["meta"])
    dcfg = DofTfmDecCfg(T, N * 3, D, 8, 128, 2)
    n = L.dof_tfm_dec_num_entries(C.byref(dcfg))
    name = C.create_string_buffer(128)
    off, numel, ndim = C.c_int64(), C.c_int64(), C.c_int()
    shape = (C.c_int * 4)()
    got = []
    for i in range(n):
        assert L.dof_tfm_dec_entry(C.byref(dcfg), i, name, C.byref(off), C.byref(numel), C.byref(ndim), shape) == 0
        got.append((name.value.decode(), tuple(shape[: ndim.value])))
    ref = [(k[len("p/decoder."):], tuple(g[k].shape)) for k in g if k.startswith("p/decoder.")]
    assert got == ref
    # bad geometry is an error code, not a crash
    bad = DofTfmCfg(T, N, E, 3, 1, D, 42, 4, 128, 2)            # key_dim not a multiple of heads
    assert L.dof_tfm_num_entries(C.byref(bad)) < 0 and L.dof_tfm_numel(C.byref(bad)) < 0
    assert "key_dim" in L.dof_last_error().decode()


def test_tcn_state_layout_is_the_reference_state_dict():
    """dof_state_entry for DOF_ENCODER_TCN lists exactly the reference's state_dict (names, order, shapes) of the three models built
    with encoder_type="TCN" (goldens from the unmodified reference), and the workspace query works without a device."""
    import ctypes as C
    from deepof_b200 import _lib
    from deepof_b200.vade import state_layout
    from helpers import load_golden_of
    for kind, case, model in (("tcnvade", "cfg2", _lib.MODEL_VADE), ("tcnvade", "main", _lib.MODEL_VADE), ("tcnstep", "vq_small", _lib.MODEL_VQVAE),
                              ("tcnstep", "con_cfg", _lib.MODEL_CONTRASTIVE)):
        g = load_golden_of(kind, case)
        T, N, E, D, K, B = (int(v) for v in g["meta"])
        cfg = _lib.DofConfig(T // 2 if model == _lib.MODEL_CONTRASTIVE else T, N, E, 3, 1, D, K, model, _lib.ENCODER_KINDS["TCN"])
        lay = state_layout(cfg)
        assert [k for k, *_ in lay] == [k[2:] for k in g if k.startswith("p/")]
        for k, off, numel, shape, grp in lay:
            assert tuple(shape) == tuple(g["p/" + k].shape), k
            assert (grp == 0) == (k.endswith(("running_mean", "running_var", "num_batches_tracked")) or k in
                                  ("encoder.laplacian", "encoder.edge_laplacian", "encoder.incidence", "latent_space.prior", "latent_space.pretrain",
                                   "latent_space.lens.weight", "latent_space.lens.bias")), k
        assert _lib.lib().dof_workspace_bytes(C.byref(cfg), 256, 1) > _lib.lib().dof_workspace_bytes(C.byref(cfg), 256, 0) > 0
