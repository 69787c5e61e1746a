"""The SOURCE of gtn_b200/csrc/k_dense.cu (factored dense-trellis Viterbi: viterbiPath / viterbiScore of
compose(emissions, transitions) without the lattice, BASELINE config 4) run on the CPU through the SIMT
emulator of tests/emu against the oracle: paths compared with ==, scores bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import emu_build

HERE = os.path.dirname(os.path.abspath(__file__))
f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def emu():
    lib = C.CDLL(emu_build.build('dense', ['k_dense.cu']))
    lib.emu_viterbi_dense.argtypes = [C.c_int, C.c_int, C.c_int, f32p, i32p, f32p, i32p, f32p]
    return lib


@pytest.mark.parametrize("shape", [(2, 9, 4), (2, 23, 8), (1, 17, 28), (2, 12, 64), (1, 10, 128)])
@pytest.mark.parametrize("ties", [False, True])
def test_dense_viterbi_kernel_source_matches_oracle(emu, oracle, shape, ties):
    B, T, Cn = shape
    rng = np.random.default_rng(Cn + 1000 * ties)
    if ties:  # small integers: many equal-score paths, the first-relaxed predecessor must win
        e = rng.integers(-2, 3, (B, T, Cn)).astype(np.float32)
        tw = rng.integers(-1, 2, Cn + Cn * Cn).astype(np.float32)
    else:
        e = rng.uniform(-5, 5, (B, T, Cn)).astype(np.float32)
        tw = rng.uniform(-5, 5, Cn + Cn * Cn).astype(np.float32)
    lens = np.array([T - 2 * b for b in range(B)], np.int32)
    paths = np.full((B, T), -7, np.int32)
    scores = np.zeros(B, np.float32)
    rc = emu.emu_viterbi_dense(B, T, Cn, e.ctypes.data_as(f32p), lens.ctypes.data_as(i32p), tw.ctypes.data_as(f32p),
                               paths.ctypes.data_as(i32p), scores.ctypes.data_as(f32p))
    emu_build.check(rc)
    for b in range(B):
        want_path, want_score = oracle.viterbi_dense(e[b, :lens[b]], tw)
        assert np.array_equal(paths[b, :lens[b]], want_path), (b, paths[b, :lens[b]], want_path)
        assert scores[b] == np.float32(want_score), (b, scores[b], want_score)


def test_emulator_skips_when_the_host_refuses_threads(emu, monkeypatch):
    """A sandbox with a low thread limit must skip the emulation tests, not abort the suite: the CTA's
    threads park until all of them exist, and the driver reports 77 when one could not be created."""
    monkeypatch.setenv("GTNB_EMU_MAX_THREADS", "100")
    e = np.zeros((1, 4, 8), np.float32)
    tw = np.zeros(8 + 64, np.float32)
    paths = np.zeros((1, 4), np.int32)
    scores = np.zeros(1, np.float32)
    rc = emu.emu_viterbi_dense(1, 4, 8, e.ctypes.data_as(f32p), None, tw.ctypes.data_as(f32p), paths.ctypes.data_as(i32p),
                               scores.ctypes.data_as(f32p))
    assert rc == 77
    with pytest.raises(pytest.skip.Exception):
        emu_build.check(rc)
    monkeypatch.delenv("GTNB_EMU_MAX_THREADS")
    rc = emu.emu_viterbi_dense(1, 4, 8, e.ctypes.data_as(f32p), None, tw.ctypes.data_as(f32p), paths.ctypes.data_as(i32p),
                               scores.ctypes.data_as(f32p))
    assert rc == 0  # and the next launch works again
