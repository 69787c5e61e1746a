"""The SOURCE of gtn_b200/csrc/k_linear.cu (the CTC normaliser: forwardScore / viterbiScore of the
emissions chain and its gradient, shortest.cpp:86-188 on creations.cpp:20-33) run on the CPU through the
SIMT emulator of tests/emu, scalar and float4 kernels, against numpy."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import emu_build

HERE = os.path.dirname(os.path.abspath(__file__))
f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def emu():
    lib = C.CDLL(emu_build.build('linear', ['k_linear.cu']))
    lib.emu_linear.argtypes = [C.c_int, C.c_int, C.c_int, f32p, i32p, C.c_int, C.c_int, C.c_float, C.c_int, f32p, f32p]
    return lib


def run(lib, e, lens, tropical, vec, delta=1.0, overwrite=1, grad0=None):
    B, T, Cn = e.shape
    e = np.ascontiguousarray(e, np.float32)
    lens = np.ascontiguousarray(lens, np.int32)
    scores = np.zeros(B, np.float32)
    grad = np.zeros((B, T, Cn), np.float32) if grad0 is None else grad0.copy()
    rc = lib.emu_linear(B, T, Cn, e.ctypes.data_as(f32p), lens.ctypes.data_as(i32p), int(tropical), int(vec), delta,
                        overwrite, scores.ctypes.data_as(f32p), grad.ctypes.data_as(f32p))
    return rc, scores, grad


@pytest.mark.parametrize("C_", [4, 8, 16, 28, 32, 64, 128, 5])
@pytest.mark.parametrize("tropical", [False, True])
def test_normaliser_kernels_match_numpy(emu, C_, tropical):
    rng = np.random.default_rng(C_ + 100 * tropical)
    B, T = 2, 37 if C_ >= 64 else 70  # T not a multiple of the rows per block; ragged below
    e = rng.uniform(-5, 5, (B, T, C_)).astype(np.float32)
    e[0, 3, :] = e[0, 3, 0]  # a row of ties: the first maximum wins (shortest.cpp:124)
    lens = np.array([T, T - 11], np.int32)
    x = e.astype(np.float64)
    mx = x.max(2, keepdims=True)
    if tropical:
        rows = mx[..., 0]
        g = np.zeros_like(x)
        am = x.argmax(2)
        for b in range(B):
            g[b, np.arange(T), am[b]] = 1.0
    else:
        rows = mx[..., 0] + np.log(np.exp(x - mx).sum(2))
        g = np.exp(x - rows[..., None])
    want_s = np.array([rows[b, :lens[b]].sum() for b in range(B)])
    vec_ok = C_ in (4, 8, 16, 32, 64, 128)
    for vec in ((0, 1) if vec_ok else (0,)):
        rc, s, grad = run(emu, e, lens, tropical, vec, delta=0.5)
        emu_build.check(rc)
        assert np.allclose(s, want_s, rtol=1e-6), (vec, s, want_s)
        for b in range(B):
            assert np.allclose(grad[b, :lens[b]], 0.5 * g[b, :lens[b]], rtol=1e-5, atol=1e-7), (vec, b)
            assert not grad[b, lens[b]:].any()
        # accumulate instead of overwrite
        base = rng.uniform(-1, 1, e.shape).astype(np.float32)
        rc, s2, grad2 = run(emu, e, lens, tropical, vec, delta=0.5, overwrite=0, grad0=base)
        for b in range(B):
            assert np.allclose(grad2[b, :lens[b]], base[b, :lens[b]] + 0.5 * g[b, :lens[b]], rtol=1e-5, atol=1e-6)
            assert np.array_equal(grad2[b, lens[b]:], base[b, lens[b]:])
    if not vec_ok:
        assert run(emu, e, lens, tropical, 1)[0] == 1


def test_normaliser_non_finite_rows(emu):
    """-inf everywhere / a +inf entry: the row score is the maximum itself and the gradient NaN
    (shortest.cpp:62-80, pinned by test/autograd_test.cpp:340-386), in both kernels."""
    e = np.zeros((1, 3, 8), np.float32)
    e[0, 1, :] = -np.inf
    e[0, 2, 3] = np.inf
    for vec in (0, 1):
        rc, s, grad = run(emu, e, np.array([3], np.int32), False, vec)
        assert np.isnan(s[0])  # -inf + inf
        assert np.allclose(grad[0, 0], 1.0 / 8)
        assert np.isnan(grad[0, 1]).all() and np.isnan(grad[0, 2]).any()
