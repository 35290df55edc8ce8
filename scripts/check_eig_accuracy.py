#!/usr/bin/env python
"""Accuracy of the device eigenvalues of the Bardeen path against LAPACK's bisection (stebz, the accurate tridiagonal
driver) and against dense eigvalsh (the reference's route), on right-well matrices of the cfg5 sweep (S = 1651).
    python scripts/check_eig_accuracy.py"""
import json, os, sys
import numpy as np
from scipy.linalg import eigh_tridiagonal
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transport_b200 import _lib, bardeen
from transport_b200._lib import as_f64, check, ptr

lib = _lib.load()
one = np.eye(1)
HC = np.zeros((11, 11, 1, 1), dtype=complex)
for j in range(11):
    HC[j, j] = 0.4
for j in range(10):
    HC[j, j + 1] = HC[j + 1, j] = -1.0
VRs = np.linspace(-0.05, 0.05, 64)
d, e = [], []
for v in VRs:
    VR = v * one
    b = bardeen.hsys_site_blocks(one, one, one, 0.5 * one, VR + 0.5 * one, VR, 20, 800, 800, HC)
    a, c = bardeen._channel_tridiagonals(b)
    d.append(a[0]); e.append(c[0])
d, e = as_f64(np.array(d)), as_f64(np.array(e))
n_mat, S = d.shape
cut = as_f64(np.full(n_mat, 0.11 - 2.0))
out = {}
for mode in (1, 2):
    check(lib.tx_set_option(3, mode))
    ns = np.zeros(n_mat, dtype=np.int32)
    check(lib.tx_tridiag_eigh_below(ptr(d), ptr(e), S, n_mat, ptr(cut), None, 0, None, None, ptr(ns), None))
    ms = int(ns.max())
    ev = np.zeros((n_mat, ms)); vec = np.zeros((n_mat, ms, S))
    check(lib.tx_tridiag_eigh_below(ptr(d), ptr(e), S, n_mat, ptr(cut), None, ms, ptr(ev), ptr(vec), ptr(ns), None))
    err_s = err_d = res = 0.0
    for i in range(0, n_mat, 7):
        k = int(ns[i])
        ws = eigh_tridiagonal(d[i], e[i], eigvals_only=True, lapack_driver="stebz", select="i", select_range=(0, k - 1))
        T = np.diag(d[i]) + np.diag(e[i], 1) + np.diag(e[i], -1)
        wd = np.linalg.eigvalsh(T)[:k]
        err_s = max(err_s, float(np.abs(ev[i, :k] - ws).max()))
        err_d = max(err_d, float(np.abs(ev[i, :k] - wd).max()))
        Z = vec[i, :k]
        res = max(res, float(np.abs(Z @ T - ev[i, :k, None] * Z).max()))
    out["warp_per_eigenvalue" if mode == 1 else "lane_per_eigenvalue"] = {
        "max_abs_err_vs_stebz": err_s, "max_abs_err_vs_dense_eigvalsh": err_d, "max_residual": res}
check(lib.tx_set_option(3, 0))
print(json.dumps(out))
