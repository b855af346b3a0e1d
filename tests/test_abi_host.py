"""CPU-side checks of the C ABI: the library loads, exports every symbol include/lcn_b200.h declares,
and its host-only entry points (mask construction, model layout) are bit exact.  No GPU compute."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from lcn_pose_b200 import _lib as L
from oracle import lcn_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "lcn_b200.h")).read()
    declared = set(re.findall(r"\b(lcn_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"lcn_model_desc"}
    lib = L.load()
    assert declared == set(L.PROTOTYPES), (declared ^ set(L.PROTOTYPES))
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.lcn_version()


@pytest.mark.parametrize("knn", [1, 2, 3, 4, 5])
def test_neighbour_matrix_bit_exact_vs_reference(golden, knn):
    assert L.neighbour_matrix(knn).tobytes() == golden[f"neighbour_knn{knn}"].tobytes()


def test_exponential_matrix_bit_exact_vs_reference(golden):
    assert L.exponential_matrix().tobytes() == golden["exponential"].tobytes()


def _create(F=64, in_F=2, L_=3, knn=3, mask_kind=0, path=1, **kw):
    d = L.ModelDesc()
    d.F, d.in_F, d.num_layers, d.mask_kind, d.path = F, in_F, L_, mask_kind, path
    d.residual, d.batch_norm, d.max_norm = kw.get("residual", 1), kw.get("batch_norm", 1), kw.get("max_norm", 1)
    sup = O.get_neighbour_matrix_by_hand(knn=knn).T
    d.support[:] = sup.reshape(-1).tolist()
    h = C.c_void_p()
    rc = L.load().lcn_model_create(C.byref(d), C.byref(h))
    return rc, h


def test_model_layout_matches_reference_variables():
    rc, h = _create()
    assert rc == 0
    lib = L.load()
    names, total = [], 0
    name = C.create_string_buffer(256)
    off, rows, cols = C.c_int64(), C.c_int32(), C.c_int32()
    for i in range(lib.lcn_model_num_tensors(h)):
        assert lib.lcn_model_tensor_info(h, i, name, 256, C.byref(off), C.byref(rows), C.byref(cols)) == 0
        names.append(name.value.decode())
        assert off.value % 4 == 0
        total += rows.value * cols.value
    cfg = O.LcnConfig(neighbour_matrix=O.get_neighbour_matrix_by_hand(knn=3))
    expect = set(O.init_params(cfg).keys())
    assert set(names) == expect
    assert total == 7203796                      # L=3, F=64 parameter count
    assert lib.lcn_model_param_count(h) >= total
    assert lib.lcn_model_workspace_bytes(h, 4096, 4096, 1) > 0
    lib.lcn_model_destroy(h)


def test_invalid_arguments_are_errors_not_ub():
    lib = L.load()
    rc, _ = _create(F=48)
    assert rc == -1 and b"F=48" in lib.lcn_last_error()
    rc, _ = _create(batch_norm=0)
    assert rc == -1
    rc, _ = _create(in_F=7)
    assert rc == -1
    out = np.zeros((17, 17), np.float32)
    assert lib.lcn_neighbour_matrix(0, out.ctypes.data) == -1
