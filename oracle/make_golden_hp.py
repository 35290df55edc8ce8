"""
TEST INFRASTRUCTURE.  Extended-precision fixtures for the ill-conditioned full-size cases (run in the build
container; minutes of CPU):

    python -m oracle.make_golden_hp cfg3      -> tests/golden/hp_cfg3_M10002.npz
    python -m oracle.make_golden_hp cfg4      -> tests/golden/hp_cfg4_leads.npz

cfg3 (BASELINE config 3 at full length, n=8, M=10 002): for the energies with the worst device unitarity residual
(gpurun_out/cfg3_worst_idx.json, written on a B200 by scripts/cfg3_worst.py; the list is stored in the fixture) plus a
uniform stride: R, T from the long-double recursion (oracle/rgf_oracle_hp.py), cross-checked (and, where it ran,
replaced) by mpmath at 50 digits on a subset that includes the four worst energies; the float64 numpy recursion's own error against that truth; and the condition number
kappa(E) = max |dT/T| / |dE/E| of the problem, measured in long double by perturbing E by 2^-40 relative.
"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import rgf_oracle, rgf_oracle_hp as hp  # noqa: E402
from transport_b200 import configs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _relerr(got, want):
    want = np.asarray(want, dtype=float)
    den = np.maximum(np.abs(want), 1e-13 * np.abs(want).max(axis=-1, keepdims=True))
    return np.abs(np.asarray(got, dtype=float) - want) / den


def _ld_chunk(args):
    h, tnn, Es = args
    src = np.eye(h.shape[-1])
    R, T = hp.transmission_closed_ld(h, tnn, Es, src)
    Rp, Tp = hp.transmission_closed_ld(h, tnn, np.real(Es).astype(np.longdouble) * (1 + np.longdouble(2.0) ** -40), src)
    den = np.maximum(np.abs(T), np.longdouble(1e-13) * np.abs(T).max(axis=-1, keepdims=True))
    kap = (np.abs(Tp - T) / den).max(axis=(1, 2)) * np.longdouble(2.0) ** 40
    unit = np.abs(R.sum(-1) + T.sum(-1) - 1).max(axis=1)
    return R.astype(float), T.astype(float), kap.astype(float), unit.astype(float)


def _mp_one(args):
    h, tnn, E = args
    R, T, w = hp.transmission_closed_mp(h, tnn, E, np.eye(h.shape[-1]), dps=50)
    return R, T, w


def cfg3():
    h, tnn, _ = configs.disordered_blocks(N=10_000, n=8, W=0.05, seed=1234)
    Es = np.linspace(-1.9, 1.9, 100_000).astype(complex)
    worst_path = os.path.join(ROOT, "gpurun_out", "cfg3_worst_idx.json")
    worst = json.load(open(worst_path))["worst_idx"][:48] if os.path.exists(worst_path) else []
    keep = set()
    old = os.path.join(OUT, "hp_cfg3_M10002.npz")
    if os.path.exists(old):                    # energies of earlier fixtures stay (they were the worst of earlier kernels)
        keep = set(int(i) for i in np.load(old)["idx"])
    idx = np.array(sorted(set(worst) | keep | set(range(1000, 100_000, 3125))), dtype=np.int64)
    print("cfg3: %d energies (%d worst-unitarity + stride)" % (len(idx), len(worst)), flush=True)
    cores = os.cpu_count() or 1
    t0 = time.time()
    with mp.get_context("fork").Pool(cores) as pool:
        parts = pool.map(_ld_chunk, [(h, tnn, Es[idx[i::cores]]) for i in range(cores)])
    R = np.zeros((len(idx), 8, 8)); T = np.zeros_like(R); kap = np.zeros(len(idx)); unit = np.zeros(len(idx))
    for i, (r, t, k, u) in enumerate(parts):
        R[i::cores], T[i::cores], kap[i::cores], unit[i::cores] = r, t, k, u
    print("long double done in %.0f s; its own unitarity residual max %.1e" % (time.time() - t0, unit.max()), flush=True)
    oR, oT = rgf_oracle.transmission_closed(h, tnn, Es[idx], np.eye(8))
    f64_err = np.maximum(_relerr(oT, T).max(axis=(1, 2)), _relerr(oR, R).max(axis=(1, 2)))
    f64_unit = np.abs(oR.sum(-1) + oT.sum(-1) - 1).max(axis=1)
    # mpmath cross-check of the truth on the 4 worst + 4 strided energies
    sub = list(dict.fromkeys([int(np.where(idx == w)[0][0]) for w in worst[:5]] + list(range(0, len(idx), max(1, len(idx) // 3)))[:3]))
    if os.environ.get("TX_HP_ALL_MP"):          # mpmath on every energy (about half an hour on 8 cores)
        sub = list(range(len(idx)))
    t0 = time.time()
    with mp.get_context("fork").Pool(min(cores, len(sub))) as pool:
        mps = pool.map(_mp_one, [(h, tnn, Es[idx[k]]) for k in sub], chunksize=1)
    ld_vs_mp = max(max(_relerr(T[k], m[1]).max(), _relerr(R[k], m[0]).max()) for k, m in zip(sub, mps))
    print("mpmath (50 digits) on %d energies in %.0f s: long double deviates by %.2e, mp unitarity %.1e"
          % (len(sub), time.time() - t0, ld_vs_mp, max(m[2] for m in mps)), flush=True)
    # the long-double truth runs the plain recursion, which loses cond(A_j) * eps_ld at a nearly singular A_j (1e8 * 1e-19
    # at the worst energy met): good to <= 1e-10 everywhere, typically 1e-13; where mpmath ran, its values are stored
    for k, m in zip(sub, mps):
        dev_k = max(_relerr(T[k], m[1]).max(), _relerr(R[k], m[0]).max())
        assert dev_k < 1e-7, (k, dev_k, kap[k])
        R[k], T[k] = m[0], m[1]
    # the float64 recursion's error against the FINAL truth (mpmath where it ran)
    f64_err = np.maximum(_relerr(oT, T).max(axis=(1, 2)), _relerr(oR, R).max(axis=(1, 2)))
    np.savez(os.path.join(OUT, "hp_cfg3_M10002.npz"), idx=idx, Es=Es[idx], R_hp=R, T_hp=T, kappa=kap, f64_err=f64_err,
             f64_unitarity=f64_unit, hp_unitarity=unit, mp_checked=np.array([idx[k] for k in sub]),
             ld_vs_mp_max_rel=ld_vs_mp, worst_idx=np.array(worst, dtype=np.int64))
    print("float64 numpy recursion vs truth: max %.2e median %.2e; kappa max %.2e; err/(kappa eps) max %.2f"
          % (f64_err.max(), np.median(f64_err), kap.max(), (f64_err / (kap * 2.2e-16)).max()))


def cfg4():
    """Sancho-Rubio self-energy of the cfg4 ribbon lead at eta = 1e-8 in long double (same decimation, tol 1e-19)."""
    n = 64
    h, tnn, _ = configs.ribbon_blocks(width=n, L=32, W=1.0, seed=1234)
    Es = np.linspace(-3.5, 3.5, 4096)
    idx = np.arange(37, 4096, 256)
    eta = np.longdouble(1e-8)
    CLD, LD = np.clongdouble, np.longdouble
    z = Es[idx].astype(LD) + 1j * eta.astype(CLD)
    eye = np.eye(n, dtype=CLD)
    h00, h01 = h[0].astype(CLD), tnn[0].astype(CLD)
    # left lead (side -1): g = (z - h00 - h01^dag g h01)^-1; decimation on alpha = h01^dag, beta = h01
    nE = len(idx)
    zE = z[:, None, None] * eye
    eps_s = np.broadcast_to(h00, (nE, n, n)).copy()
    eps = eps_s.copy()
    alpha = np.broadcast_to(np.conj(h01).T, (nE, n, n)).copy()
    beta = np.broadcast_to(h01, (nE, n, n)).copy()
    for it in range(80):
        g = hp._inv_ld(zE - eps)
        ag, bg = np.matmul(alpha, g), np.matmul(beta, g)
        agb, bga = np.matmul(ag, beta), np.matmul(bg, alpha)
        eps_s = eps_s + agb
        eps = eps + agb + bga
        alpha, beta = np.matmul(ag, alpha), np.matmul(bg, beta)
        if max(np.abs(alpha).max(), np.abs(beta).max()) < 1e-19:
            break
    gs = hp._inv_ld(zE - eps_s)
    sig = np.matmul(np.matmul(np.conj(h01).T, gs), h01)
    res = gs - hp._inv_ld(zE - h00 - sig)
    print("cfg4 leads: %d iterations, fixed-point residual %.1e" % (it + 1, float(np.abs(res).max())))
    zs = Es[idx] + 1j * 1e-8
    g64, _ = rgf_oracle.sancho_rubio(h[0], tnn[0], zs, -1)
    s64 = np.conj(tnn[0]).T @ g64 @ tnn[0]
    den = np.abs(sig).max(axis=(1, 2), keepdims=True).astype(float)
    f64_err = (np.abs(s64 - sig.astype(complex)) / den).max(axis=(1, 2))
    np.savez(os.path.join(OUT, "hp_cfg4_leads.npz"), idx=idx, Es=Es[idx], sigL_hp=sig.astype(complex), f64_err=f64_err,
             eta=1e-8, residual=float(np.abs(res).max()))
    print("float64 numpy Sancho-Rubio vs long double: max %.2e" % f64_err.max())


if __name__ == "__main__":
    for name in sys.argv[1:] or ["cfg3", "cfg4"]:
        {"cfg3": cfg3, "cfg4": cfg4}[name]()
