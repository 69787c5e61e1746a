"""The SOURCE of gtn_b200/csrc/k_banded.cu (experimental temporally blocked CTC sweeps: K frames per
barrier, neighbour exchange through warp shuffles) compiled with g++ against the SIMT emulator of
tests/emu/simt_emu.h and run on the CPU, one std::thread per CUDA thread, against the oracle.

This is what can be checked without a GPU: the lane / halo arithmetic, the register prefetch blocks and
their tails, barrier placement (a wrong one deadlocks or corrupts the exchange rows), ragged lengths.
The SFU instructions are replaced by exp2f / log2f and nothing is said about speed."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import emu_build, util

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    lib = C.CDLL(emu_build.build('banded', ['k_banded.cu']))
    f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    lib.emu_banded_ctc.argtypes = [C.c_int, C.c_int, C.c_int, f32p, i32p, i32p, i32p, C.c_int, C.c_int, f32p, f32p, i32p]
    return lib


def run(lib, e, targets, lens, blank=0, K=4):
    B, T, Cn = e.shape
    e = np.ascontiguousarray(e, np.float32)
    cat = np.ascontiguousarray(np.concatenate(targets) if sum(map(len, targets)) else np.zeros(0), np.int32)
    tl = np.array([len(t) for t in targets], np.int32)
    lens = np.ascontiguousarray(lens, np.int32)
    out = np.zeros(B, np.float32)
    grad = np.zeros((B, T, Cn), np.float32)
    status = np.zeros(B, np.int32)
    f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    rc = lib.emu_banded_ctc(B, T, Cn, e.ctypes.data_as(f32p), lens.ctypes.data_as(i32p), cat.ctypes.data_as(i32p),
                            tl.ctypes.data_as(i32p), blank, K, out.ctypes.data_as(f32p), grad.ctypes.data_as(f32p),
                            status.ctypes.data_as(i32p))
    emu_build.check(rc)
    return out, grad, status


# (B, T, C, U): one warp; several warps (forward: 24 own nodes per warp, backward: 22); T below / at / above
# the 8-frame prefetch block and the 4-frame barrier interval; T == 2U+1 (almost everything pruned);
# the benchmark's graph (U = 100: 201 nodes, 9 / 10 warps)
SHAPES = [(2, 5, 4, 1), (3, 13, 6, 3), (2, 9, 5, 4), (2, 31, 8, 12), (2, 40, 16, 18), (1, 57, 28, 26),
          (1, 24, 5, 11), (1, 70, 9, 34), (1, 230, 64, 100)]


@pytest.mark.parametrize("K", [4, 1, 2, 8])
@pytest.mark.parametrize("shape", SHAPES)
def test_banded_kernel_source_matches_oracle(emu, oracle, shape, K):
    B, T, Cn, U = shape
    if K != 4 and shape not in SHAPES[2:6] + SHAPES[-1:]:
        pytest.skip("the other barrier intervals run a subset of the shapes")
    e, targets = util.bench_inputs(B, T, Cn, U, seed=1357)
    if U > 2:
        for t in targets:
            t[1] = t[0]  # a repeated label: no skip arc
    lens = np.array([max(T - 5 * b, 2 * U) for b in range(B)], np.int32)  # ragged, still feasible
    lens = np.minimum(lens, T)
    out, grad, status = run(emu, e, targets, lens, K=K)
    assert not status.any(), status
    for b in range(B):
        Tb = int(lens[b])
        lo, go = oracle.ctc_loss(e[b, :Tb], targets[b], 0, True)
        x = e[b, :Tb].astype(np.float64)
        mx = x.max(1, keepdims=True)
        lse = mx[:, 0] + np.log(np.exp(x - mx).sum(1))
        loss = float(lse.sum() - np.float64(out[b]))
        if not np.isfinite(lo):
            assert not np.isfinite(out[b])
            continue
        assert util.close(loss, lo), (b, loss, lo)
        g = np.exp(x - lse[:, None]) + grad[b, :Tb]  # normaliser's softmax + (-1) x lattice posterior
        assert util.grad_close(g, go, 5.0 * T), (b, float(np.abs(g - go).max()))
        assert not grad[b, Tb:].any()
        # the lattice posterior sums to one per frame
        assert np.allclose(-grad[b, :Tb].sum(1), 1.0, atol=1e-3)


def test_banded_kernel_reports_non_finite_weights(emu):
    B, T, Cn, U = 1, 12, 5, 3
    e, targets = util.bench_inputs(B, T, Cn, U, seed=5)
    e[0, 4, int(targets[0][1])] = -np.inf
    out, grad, status = run(emu, e, targets, np.array([T], np.int32))
    assert status[0] & 1
