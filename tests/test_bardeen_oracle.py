"""CPU tests: oracle/bardeen_oracle.py against tests/golden/bardeen_*.npz, the outputs of the unmodified
reference's transport.bardeen at the driver parameters of BASELINE config 5 (oracle/make_golden_bardeen.py)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import bardeen_oracle as bo

WELL = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "bardeen_*.npz")))


def _run(g):
    common = (g["tinfty"], g["tL"], g["tR"], g["Vinfty"], g["VL"], g["VLprime"], g["VR"], g["VRprime"],
              int(g["Ninfty"]), int(g["NL"]), int(g["NR"]), g["HC"])
    if "HCprime" in g.files:
        return bo.kernel_well_prime(*common, g["HCprime"], g["E_cutoff"], interval=float(g["interval"]))
    return bo.kernel_well(*common, g["HCobs"], g["defines_Sz"], g["E_cutoff"], interval=float(g["interval"]))


def m2_err(M, Mg, floor=1e-30):
    """|M - Mg| / Mg on the entries above `floor` (entries below are eigensolver noise in the reference:
    bound states with no partner inside the energy window), absolute `floor` on the rest."""
    M, Mg = np.asarray(M), np.asarray(Mg)
    assert M.shape == Mg.shape
    big = Mg > floor
    assert np.all(np.abs(M[~big]) <= floor)
    return float(np.max(np.abs(M[big] - Mg[big]) / Mg[big])) if big.any() else 0.0


@pytest.mark.parametrize("name", WELL)
def test_kernel_well_goldens(name):
    g = load_golden(name)
    E, M = _run(g)
    assert E.shape == g["E"].shape and np.max(np.abs(E - g["E"])) < 1e-13
    assert m2_err(M, g["M"]) < 1e-9
    if "T" in g.files:
        T = bo.Ts_bardeen(E, M, g["tL"], g["tR"], g["VL"], g["VR"], int(g["NL"]), int(g["NR"]))
        assert m2_err(T, g["T"]) < 1e-9


@pytest.mark.parametrize("name", [w for w in WELL if w.startswith("bardeen_prime_J")])
def test_current_goldens(name):
    g = load_golden(name)
    for key, muR, kBT in (("I_mu0_kT1em4", 0.0, 0.0001), ("I_mum18_kT1em2", -1.8, 0.01)):
        I = bo.current(g["E"], g["M"], g["Vb"], g["tL"], muR, kBT)
        assert np.allclose(I, g[key], rtol=1e-12, atol=1e-30)


def test_survey_spot_value():
    # kernel_well n=1, N_LR=50 -> T[0,:3,0]
    g = load_golden("bardeen_barrier_N50.npz")
    assert np.allclose(g["T"][0, :3, 0], [1.61123972e-07, 7.41202481e-07, 2.11216594e-06], rtol=1e-8)
