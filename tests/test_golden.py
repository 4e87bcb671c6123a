"""Golden vectors (tests/golden/*.npz, see make_golden.py): the oracle must keep reproducing them on
CPU; the CUDA engine must reproduce them on the GPU."""
import glob
import os

import numpy as np
import pytest

from oracle import nl_oracle as O
from tests import util as U

FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def test_golden_files_present():
    assert len(FILES) >= 6


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    dt = g["X"].dtype
    r = O.sortbased(g["X"], g["cutoff"], g["cell"], tuple(g["pbc"]), dtype=dt)
    assert np.array_equal(r["first"], g["first"]) and np.array_equal(r["perm"], g["perm"])
    assert np.array_equal(r["cell_id"], g["cell_id"]) and np.array_equal(r["cell_offsets"], g["cell_offsets"])
    i, j, S, R = O.canonical(r["i"], r["j"], r["S"], r["R"])
    assert np.array_equal(i, g["i"]) and np.array_equal(j, g["j"]) and np.array_equal(S, g["S"]) and np.array_equal(R, g["R"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_engine_reproduces_golden(path):
    import torch
    import neighbourlists_jl_b200 as nl
    g = np.load(path)
    dt = g["X"].dtype
    clist = nl.build_cell_list(torch.from_numpy(g["X"]).cuda(), g["cutoff"], g["cell"], tuple(g["pbc"]))
    pl = nl.materialize_pairlist(clist, with_R=True)
    assert np.array_equal(clist.perm.cpu().numpy(), g["perm"]) and np.array_equal(clist.cell_offsets.cpu().numpy(), g["cell_offsets"])
    gold = dict(first=g["first"], i=g["i"], j=g["j"], S=g["S"], R=g["R"])
    U.assert_engine_matches_oracle(pl.cpu(), gold, 1e-12 if dt == np.float64 else 1e-5, msg=os.path.basename(path))
