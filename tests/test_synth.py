"""The bench's matrix generator (matfac_b200/synth.py: skewed_problem) is a pure function of (shape, seed): the device arm,
the reference arm, every N and every rank must train on the same matrix (`same_config`)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matfac_b200 import synth  # noqa: E402

SMALL = (3000, 400, 120_000, 20260102)
# CRC-32 of (train rowptr, rowind, rowval, val rowptr, rowind, rowval) for SMALL, as first generated (CPU, this image)
SMALL_CRC = "5b5e686a"


def test_generator_is_reproducible_and_pinned():
    a = synth.skewed_problem(*SMALL)
    b = synth.skewed_problem(*SMALL)
    assert a["crc"] == b["crc"] == synth.array_crc(*a["train"], *a["val"])
    assert a["crc"] == SMALL_CRC, a["crc"]
    c = synth.skewed_problem(SMALL[0], SMALL[1], SMALL[2], SMALL[3] + 1)
    assert c["crc"] != a["crc"]


def test_generator_shape_properties():
    nu, ni, nnz, seed = SMALL
    p = synth.skewed_problem(nu, ni, nnz, seed)
    ptr, ind, val = p["train"]
    vptr, vind, vval = p["val"]
    assert ptr.shape[0] == nu + 1 and vptr.shape[0] == nu + 1          # every file has n_users rows (model.cpp:223)
    assert abs(int(ptr[-1]) + int(vptr[-1]) - int(nnz * 1.01)) <= 1
    assert np.diff(ptr).min() >= 1                                     # every user rated (io.cpp:742-752)
    assert np.bincount(ind, minlength=ni).min() >= 1                   # every item rated
    inner = np.ones(ind.shape[0], bool)
    inner[ptr[1:-1][ptr[1:-1] < ind.shape[0]]] = False
    assert np.all(np.diff(ind)[inner[1:]] > 0)                         # items ascend inside a row, no duplicates (util.cpp:919)
    assert set(np.unique(val)).issubset({1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0, 4.5, 5.0})
    top = np.bincount(ind, minlength=ni).max() / ptr[-1]
    assert top < 0.05                                                  # head of the item distribution is capped


@pytest.mark.gpu
def test_generator_is_device_independent():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    a = synth.skewed_problem(*SMALL, device="cpu")
    b = synth.skewed_problem(*SMALL, device="cuda:0")
    assert a["crc"] == b["crc"]
    for x, y in zip(a["train"] + a["val"], b["train"] + b["val"]):
        assert np.array_equal(x, y)
