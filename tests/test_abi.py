"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/rsvdt.h declares, its pure-host
entry points agree with the oracle, and compute entry points fail loudly (no CPU fallback) without a GPU."""
import ctypes as C

import pytest
import torch


def test_library_exports_every_declared_symbol(pkg):
    pkg.capi.build()
    L = pkg.capi.lib()
    names = pkg.capi.exported_symbols_from_header()
    assert len(names) >= 25
    for n in names:
        assert getattr(L, n) is not None
    assert b"sm_100a" in L.rsvdt_version()


def test_grid_arithmetic_matches_oracle(pkg, O):
    for args in ((0.05, 42, 0.003, 10.0), (0.05, 64, 0.003, 10.0), (0.8, 6, 0.004, 2.0), (2 * 3.14159, 10, 0.01, 0.0)):
        assert pkg.capi.setup_time_freq_grid(*args) == O.setup_time_freq_grid(*args)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(pkg):
    with pytest.raises(pkg.capi.RsvdtError):
        pkg.capi.RSVDt(0)


def test_synthetic_operators_are_stable(pkg):
    import numpy as np
    N, rp, ci, va = pkg.synthetic.ginzburg_landau(500)
    assert rp[-1] == 4 * N
    A = np.zeros((N, N), complex)
    rows = np.repeat(np.arange(N), 4)
    np.add.at(A, (rows, ci), va)
    ev = np.linalg.eigvals(A)
    assert ev.real.max() < 0 and np.abs(ev).max() * 0.003 < 1.0
    N, rp, ci, va = pkg.synthetic.lns_stencil((6, 7), nvar=5)
    assert N == 210 and np.diff(rp).max() == 45
    A = np.zeros((N, N), complex)
    np.add.at(A, (np.repeat(np.arange(N), np.diff(rp)), ci), va)
    assert np.linalg.eigvals(A).real.max() < 0
    N3, rp3, ci3, va3 = pkg.synthetic.lns_stencil((4, 4, 5), nvar=5)
    assert np.diff(rp3).max() == 135
    # row_range builds the same rows
    N2, rp2, ci2, va2 = pkg.synthetic.lns_stencil((6, 7), nvar=5, row_range=(50, 120))
    assert np.array_equal(ci2, ci[rp[50]:rp[120]]) and np.array_equal(va2, va[rp[50]:rp[120]])
