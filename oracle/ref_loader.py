"""
TEST INFRASTRUCTURE — not product code.

Loader for the *unmodified* reference (cpbunker/transport) from /root/reference.
Only usable in the build container (the GPU box has no /root/reference); used by
`oracle/make_golden.py` to generate the committed fixtures under tests/golden/ and
by the (auto-skipping) live cross-check in tests/test_oracle.py.

Recipe: `transport/wfm/__init__.py:11-12` imports `transport.tdfci`,
whose `__init__` does `from pyscf import lib, fci, scf, gto, ao2mo`
(`transport/tdfci/__init__.py:28-29`).  pyscf is absent, and wfm only uses the pure
numpy helper `tdfci.utils.mat_2d_to_4d`, so empty module stubs are registered for the
pyscf names before import.  `transport.bardeen` imports `transport.fci_mod`
(`transport/bardeen/__init__.py:9`), a module that no longer exists; it is aliased to
`transport.tdfci.utils`.  Driver scripts (for their Hamiltonian builders) additionally
need matplotlib stubs.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TRANSPORT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "transport", "wfm", "__init__.py"))


def _stub(name: str) -> types.ModuleType:
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__path__ = []  # behave like a package so `from a.b import c` resolves
    sys.modules[name] = mod
    if "." in name:
        parent, child = name.rsplit(".", 1)
        setattr(_stub(parent), child, mod)
    return mod


def load_reference(with_bardeen: bool = False, with_scripts: bool = False):
    """Return the reference's `transport.wfm` module (and optionally more).

    Returns a dict with keys 'wfm', 'fci_mod' and optionally 'bardeen', 'run_wfm_gate',
    'run_wfm_single'.
    """
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    for name in ("pyscf", "pyscf.lib", "pyscf.fci", "pyscf.scf", "pyscf.gto", "pyscf.ao2mo",
                 "pyscf.fci.direct_uhf", "pyscf.fci.cistring", "pyscf.fci.direct_spin1",
                 "pyscf.mcscf", "pyscf.tools"):
        _stub(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    out = {}
    out["wfm"] = importlib.import_module("transport.wfm")
    out["fci_mod"] = importlib.import_module("transport.tdfci.utils")
    if with_bardeen:
        sys.modules.setdefault("transport.fci_mod", out["fci_mod"])
        import transport  # noqa
        transport.fci_mod = out["fci_mod"]
        for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm"):
            _stub(name)
        out["bardeen"] = importlib.import_module("transport.bardeen")
    if with_scripts:
        for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors",
                     "matplotlib.ticker"):
            _stub(name)
        plt = sys.modules["matplotlib.pyplot"]
        # scripts touch rcParams / style at import time
        plt.rcParams = {}
        plt.style = types.SimpleNamespace(use=lambda *a, **k: None)
        plt.rcParams_update = lambda *a, **k: None
        argv = sys.argv
        sys.argv = [argv[0]]
        try:
            for script in ("run_wfm_gate", "run_wfm_single"):
                try:
                    out[script] = importlib.import_module(script)
                except SystemExit:
                    pass
                except Exception as exc:  # scripts run plotting code at import when argv present
                    out[script + "_error"] = exc
        finally:
            sys.argv = argv
    return out
