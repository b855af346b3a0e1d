"""Pins the model / trainer restatement (oracle/lcn_oracle.py) to the committed vectors tests/golden/model_oracle.npz
(written by tests/golden/make_model_golden.py): any drift of the oracle -- the checker of every GPU parity test --
fails here, on CPU.  The vectors are outputs of the restatement itself (TensorFlow is not installable: the model half of
the oracle stays "parity unpinned" against the real reference)."""
import os

import numpy as np
import pytest

from tests.golden import make_model_golden as G

ROOT = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def vectors():
    return np.load(os.path.join(ROOT, "golden", "model_oracle.npz"))


@pytest.mark.parametrize("case", sorted(G.CASES))
def test_oracle_reproduces_committed_model_vectors(vectors, case):
    got = G.run_case(G.CASES[case])
    keys = [k for k in vectors.files if k.startswith(case + "/")]
    assert len(keys) == len(got) and len(keys) > 10
    for k in keys:
        want, have = vectors[k], got[k[len(case) + 1:]]
        scale = max(float(np.abs(want).max()), 1e-30)
        assert np.abs(have - want).max() <= 1e-9 * scale, k
