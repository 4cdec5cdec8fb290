import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return entry.load_package()


@pytest.fixture(scope="session")
def O():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def A_GL(pkg, O):
    """the tutorial operator shipped by the reference (Tutorial/Ginzburg-Landau/A_GL), copied as a data fixture"""
    M, N, rp, ci, va = pkg.petsc_io.read_mat(os.path.join(GOLDEN, "A_GL"))
    assert (M, N) == (500, 500)
    return O.Csr(M, N, rp, ci, va)


def make_vars(O, pkg, N, k, Nw, w, dt_in, TL, real=False, TR=False, direct=True):
    dt, Ns, Nt, RR = O.setup_time_freq_grid(w, Nw, dt_in, TL)
    return O.Vars(N=N, k=k, Nw=Nw, Nw_eff=Nw // 2 if real else Nw, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=w,
                  RealOperator=real, TransientRemoval=TR, DirAdj=direct)


def rel_err(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def principal_angle(U, V):
    """largest principal angle between span(U) and span(V).  Small angles are taken from the sine -- the norm of the part of
    span(V) outside span(U) -- because arccos of a cosine within 1e-16 of one cannot resolve anything below 2e-8."""
    Qu, _ = np.linalg.qr(U)
    Qv, _ = np.linalg.qr(V)
    s = np.linalg.svd(Qu.conj().T @ Qv, compute_uv=False)
    cos = float(np.clip(s.min(), -1.0, 1.0))
    if cos < 0.9:
        return float(np.arccos(cos))
    sin = float(np.linalg.norm(Qv - Qu @ (Qu.conj().T @ Qv), 2))
    return float(np.arcsin(min(sin, 1.0)))
