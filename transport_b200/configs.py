"""
Synthetic workloads for the T(E) hot path: the named geometries of BASELINE.json.

These are host-side builders of the block arrays (h[M,n,n], tnn[M-1,n,n], tnnn[M-2,n,n]) that
the reference's driver scripts construct before looping over `wfm.kernel`.  The geometry
parameters are those of the reference scripts (cited per function); the code is written
fresh on numpy Kronecker products.
"""
import numpy as np


# ----------------------------------------------------------------------------------------
# cfg1: spinless rectangular barrier
# ----------------------------------------------------------------------------------------
def barrier_blocks(Vb=0.4, NC=11, tl=1.0, n=1):
    """Rectangular barrier of NC sites of height Vb between two clean leads.

    Geometry of `rbbarrier.py:66-74` after the mapping in `transport/bardeen/__init__.py:650-669`
    (VL=VR=0), equivalently `run_wfm_barrier.py:40-60`:  h = [0, Vb x NC, 0], tnn = -tl.
    """
    M = NC + 2
    eye = np.eye(n)
    h = np.zeros((M, n, n), dtype=complex)
    h[1:-1] = Vb * eye
    tnn = np.tile(-tl * eye.astype(complex), (M - 1, 1, 1))
    tnnn = np.zeros((M - 2, n, n), dtype=complex)
    return h, tnn, tnnn


def barrier_energies(n_E=10_000, tl=1.0, logK=(-4, 0)):
    """E = K - 2 tl, K = logspace (`run_wfm_barrier.py:63-64,83`)."""
    return np.logspace(logK[0], logK[1], n_E).astype(complex) - 2 * tl


# ----------------------------------------------------------------------------------------
# cfg2: electron scattering off two localized spins (Ciccarello geometry)
# ----------------------------------------------------------------------------------------
def _spin_ops(TwoS):
    s = 0.5 * TwoS
    m = s - np.arange(TwoS + 1)                      # s, s-1, ..., -s
    Sz = np.diag(m)
    Sp = np.diag(np.sqrt(s * (s + 1) - m[1:] * (m[1:] + 1)), k=1)
    Sm = np.diag(np.sqrt(s * (s + 1) - m[:-1] * (m[:-1] - 1)), k=-1)
    return Sz, Sp, Sm


def cicc_exchange_blocks(TwoS=1):
    """(S_e.S_1, S_e.S_2) as dense matrices in the basis |electron spin> (x) |S_1> (x) |S_2>.

    Same operators as `run_wfm_gate.py:h_cicc:31-63` (J factored out): S_e.S = s_z S_z + (s_+ S_- + s_- S_+)/2.
    """
    Sz, Sp, Sm = _spin_ops(TwoS)
    I = np.eye(TwoS + 1)
    sz = np.diag([0.5, -0.5])
    sp = np.array([[0.0, 1.0], [0.0, 0.0]])
    sm = sp.T
    k3 = lambda a, b, c: np.kron(a, np.kron(b, c))
    S1 = k3(sz, Sz, I) + 0.5 * (k3(sp, Sm, I) + k3(sm, Sp, I))
    S2 = k3(sz, I, Sz) + 0.5 * (k3(sp, I, Sm) + k3(sm, I, Sp))
    return S1.astype(complex), S2.astype(complex)


def cicc_blocks(TwoS=1, tl=1.0, J=-0.5, Vq=0.0, VB=0.0, Vend=0.0, NB=3, dist=2, offset=0):
    """Block arrays of `run_wfm_gate.py:get_hblocks:80-143` (no lead-junction barriers).

    Sites: 0 = left-lead cell, 2 and 2+dist carry the two spins, then NB barrier sites at VB,
    the last block is set to Vend*I (`:129`).  n = 2 (TwoS+1)^2, M = 3 + dist + NB.
    """
    n = 2 * (TwoS + 1) ** 2
    NC = 3 + dist
    if not NB > offset:
        raise ValueError("NB must exceed offset")
    if dist <= 0:
        raise ValueError("dist must be positive")
    M = NC + NB
    S1, S2 = cicc_exchange_blocks(TwoS)
    eye = np.eye(n)
    h = np.zeros((M, n, n), dtype=complex)
    h[offset + 2] += J * S1 + Vq * eye
    h[offset + 2 + dist] += J * S2 + Vq * eye
    h[offset + NC:] += VB * eye
    h[-1] = Vend * eye
    tnn = np.tile((-tl * eye).astype(complex), (M - 1, 1, 1))
    tnnn = np.zeros((M - 2, n, n), dtype=complex)
    return h, tnn, tnnn


def cicc_affine(TwoS=1, tl=1.0, Vq=0.0, VB=0.0, Vend=0.0, NB=3, dist=2, offset=0):
    """(h0, h1, tnn, tnnn) with h(J) = h0 + J*h1 — the affine form of `cicc_blocks`."""
    h0, tnn, tnnn = cicc_blocks(TwoS, tl, 0.0, Vq, VB, Vend, NB, dist, offset)
    hJ, _, _ = cicc_blocks(TwoS, tl, 1.0, Vq, VB, Vend, NB, dist, offset)
    return h0, hJ - h0, tnn, tnnn


def cfg2_grid(n_J=1000, n_E=1000, tl=1.0):
    """J = linspace(-1, -0.001), E = logspace(-4,0) - 2 tl."""
    Js = np.linspace(-1.0, -0.001, n_J)
    Es = np.logspace(-4, 0, n_E).astype(complex) - 2 * tl
    return Js, Es


# ----------------------------------------------------------------------------------------
# single spin-1/2 impurity (Menezes), monatomic and Rice-Mele leads
# ----------------------------------------------------------------------------------------
def kondo_block(J, trunc=None, unit_cell=1):
    """(J/4)(sigma.sigma) in the basis |uu>,|ud>,|du>,|dd> (`run_wfm_single.py:h_kondo:23-50`)."""
    hs = np.zeros((4, 4), dtype=complex)
    hs[0, 0] = hs[3, 3] = 1
    hs[1, 1] = hs[2, 2] = -1
    hs[1, 2] = hs[2, 1] = 2
    hs *= J / 4
    if trunc is not None:
        hs = hs[trunc[0]:trunc[1], trunc[0]:trunc[1]]
    ns = len(hs)
    out = np.zeros((unit_cell * ns, unit_cell * ns), dtype=complex)
    out[:ns, :ns] = hs
    return out


def kondo_blocks(J=-0.5, tl=1.0):
    """n=2 (|e up, imp dw>, |e dw, imp up>), M=3: one impurity site between clean leads
    (`run_wfm_single.py` monatomic case; `rbmenez`-style single impurity)."""
    hSR = kondo_block(J, (1, 3), 1)
    h = np.array([np.zeros_like(hSR), hSR, np.zeros_like(hSR)])
    tnn = np.tile((-tl * np.eye(2)).astype(complex), (2, 1, 1))
    tnnn = np.zeros((1, 2, 2), dtype=complex)
    return h, tnn, tnnn


def ricemele_kondo_blocks(J=-0.5, u=0.0, v=-1.0, w=-1.0):
    """Rice-Mele (diatomic) leads with a Kondo impurity on the A orbital of the centre cell
    (`run_wfm_single.py:82-160`): n = 4 = {A,B} x {|e up,imp dw>, |e dw,imp up>}, M = 3."""
    diag = np.array([[+u, 0, v, 0], [0, +u, 0, v], [v, 0, -u, 0], [0, v, 0, -u]], dtype=complex)
    off = np.zeros((4, 4), dtype=complex)
    off[2, 0] = off[3, 1] = w
    hSR = kondo_block(J, (1, 3), 2) + diag
    h = np.array([diag, hSR, diag])
    tnn = np.array([off, off])
    tnnn = np.zeros((1, 4, 4), dtype=complex)
    return h, tnn, tnnn


# ----------------------------------------------------------------------------------------
# cfg3: long disordered region
# ----------------------------------------------------------------------------------------
def disordered_blocks(N=10_000, n=8, W=0.05, tl=1.0, seed=1234, hop_noise=0.0):
    """h_j = W (A_j + A_j^dag)/2 with A_j ~ N(0,1)+iN(0,1); lead cells clean; tnn = -tl I
    (optionally + hop_noise * B_j for a general-hopping variant).  SURVEY §8d cfg3."""
    rng = np.random.default_rng(seed)
    M = N + 2
    A = rng.standard_normal((N, n, n)) + 1j * rng.standard_normal((N, n, n))
    h = np.zeros((M, n, n), dtype=complex)
    h[1:-1] = W * (A + np.conj(np.swapaxes(A, 1, 2))) / 2
    tnn = np.tile((-tl * np.eye(n)).astype(complex), (M - 1, 1, 1))
    if hop_noise:
        B = rng.standard_normal((M - 1, n, n)) + 1j * rng.standard_normal((M - 1, n, n))
        B[0] = 0
        B[-1] = 0
        tnn = tnn + hop_noise * B
    tnnn = np.zeros((M - 2, n, n), dtype=complex)
    return h, tnn, tnnn


# ----------------------------------------------------------------------------------------
# cfg4: multi-channel ribbon (square lattice strip)
# ----------------------------------------------------------------------------------------
def ribbon_blocks(width=64, L=32, W=1.0, t=1.0, seed=1234):
    """Square-lattice ribbon: h00 = -t tridiag(1) across `width` transverse sites, h01 = -t I;
    L disordered slices with on-site U(-W/2, W/2) between clean lead cells.  SURVEY §8d cfg4."""
    rng = np.random.default_rng(seed)
    h00 = np.zeros((width, width), dtype=complex)
    idx = np.arange(width - 1)
    h00[idx, idx + 1] = -t
    h00[idx + 1, idx] = -t
    M = L + 2
    h = np.tile(h00, (M, 1, 1))
    dis = rng.uniform(-W / 2, W / 2, size=(L, width))
    for j in range(L):
        h[1 + j] = h[1 + j] + np.diag(dis[j])
    tnn = np.tile((-t * np.eye(width)).astype(complex), (M - 1, 1, 1))
    tnnn = np.zeros((M - 2, width, width), dtype=complex)
    return h, tnn, tnnn
