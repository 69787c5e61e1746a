"""numpy model of the banded, temporally blocked CTC sweeps of k_banded.cu (lane-exact: one array slot
per lane, shuffles as shifted reads inside a warp), checked against the plain per-frame recursion.

    python scripts/banded_model.py            # runs the checks below on random CTC problems

Why: a graph operand whose in-arcs all come from nodes u, u-1, u-2 (CTC, forced alignment) lets a warp
advance K frames without talking to the other warps if it carries 2K halo lanes on the side the
dependencies come from -- the halo lanes recompute the neighbour warp's nodes and go stale two lanes
per frame.  The CTA-wide barrier + shared-memory exchange then happens every K frames instead of
every frame, and the per-frame neighbour reads are warp shuffles.  This file pins the index
arithmetic (which lane holds which node, which lanes are still valid after j frames, what is
published / re-read at a block boundary) before it is written in CUDA.
"""
import numpy as np

NINF = -np.inf


def lse3(a, b, c):
    mx = np.maximum(np.maximum(a, b), c)
    with np.errstate(invalid="ignore"):
        s = np.exp(a - mx) + np.exp(b - mx) + np.exp(c - mx)
        r = mx + np.log(s)
    return np.where(np.isinf(mx), mx, r)


def ctc_band(target, blank=0):
    """Band form of the CTC graph (benchmarks/ctc.cpp:40-58): w[d][u] = weight of the arc u-d -> u
    (0.0) or -inf when absent; label per node; start / accept flags."""
    U = len(target)
    N = 2 * U + 1
    w = np.full((3, N), NINF)
    lab = np.zeros(N, np.int64)
    for u in range(N):
        lab[u] = target[(u - 1) // 2] if u % 2 else blank
        w[0, u] = 0.0
        if u > 0:
            w[1, u] = 0.0
        if u % 2 and u > 1 and target[(u - 1) // 2] != target[(u - 3) // 2]:
            w[2, u] = 0.0
    start = np.zeros(N, bool)
    start[0] = True
    accept = np.zeros(N, bool)
    accept[N - 1] = True
    if N > 1:
        accept[N - 2] = True
    return w, lab, start, accept


def forward_plain(e, w, lab, start):
    T = e.shape[0]
    N = w.shape[1]
    S = np.full((T + 1, N), NINF)
    S[0, start] = 0.0
    for t in range(1, T + 1):
        p = S[t - 1]
        p1 = np.concatenate([[NINF], p[:-1]])
        p2 = np.concatenate([[NINF, NINF], p[:-2]])
        em = e[t - 1, lab]
        S[t] = lse3(p + (w[0] + em), p1 + (w[1] + em), p2 + (w[2] + em))
    return S


def shfl_up(x, d):
    """per warp (rows): lane l reads lane l-d; lanes < d keep their own value (CUDA semantics)"""
    y = x.copy()
    y[:, d:] = x[:, :-d]
    return y


def shfl_down(x, d):
    y = x.copy()
    y[:, :-d] = x[:, d:]
    return y


def forward_banded(e, w, lab, start, K):
    """Forward sweep, K frames per barrier.  Lane l of warp q holds node q*own + l - H, H = 2K."""
    T = e.shape[0]
    N = w.shape[1]
    H = 2 * K
    own = 32 - H
    W = -(-N // own)
    node = (np.arange(W)[:, None] * own + np.arange(32)[None, :] - H)  # [W][32]
    ok = (node >= 0) & (node < N)
    nc = np.clip(node, 0, N - 1)
    is_own = np.arange(32)[None, :] >= H
    wl = np.where(ok[None], w[:, nc], NINF)  # [3][W][32]
    labl = lab[nc]
    S = np.full((T + 1, N), NINF)  # what the kernel stores to HBM
    s = np.where(ok & start[nc], 0.0, NINF)
    row = np.full(N, NINF)  # the shared-memory exchange row

    def publish():
        m = ok & is_own
        row[node[m]] = s[m]
        S_t[node[m]] = s[m]

    S_t = S[0]
    publish()
    t0 = 0
    while t0 < T:
        # block boundary (after the barrier): halo lanes re-read their nodes from the exchange row
        halo = ok & ~is_own
        s = np.where(halo, row[nc], np.where(is_own, s, NINF))
        for j in range(min(K, T - t0)):
            f = t0 + j
            em = e[f][labl]
            s1, s2 = shfl_up(s, 1), shfl_up(s, 2)
            s = lse3(s + (wl[0] + em), s1 + (wl[1] + em), s2 + (wl[2] + em))
            m = ok & is_own
            S[f + 1][node[m]] = s[m]  # own lanes store every frame (the backward sweep needs them)
        S_t = S[min(t0 + K, T)]
        m = ok & is_own
        row[node[m]] = s[m]
        t0 += K
    return S


def backward_plain(e, w, lab, accept, S, delta=1.0):
    T, C = e.shape
    N = w.shape[1]
    grad = np.zeros((T, C))
    sa = np.where(accept, S[T], NINF)
    mx = sa.max()
    out = mx + np.log(np.exp(sa - mx).sum())
    G = np.where(accept, np.exp(S[T] - out), 0.0)
    for t in range(T, 0, -1):
        p = S[t - 1]
        p1 = np.concatenate([[NINF], p[:-1]])
        p2 = np.concatenate([[NINF, NINF], p[:-2]])
        em = e[t - 1, lab]
        se = np.where(np.isinf(S[t]), np.inf, S[t])
        E = [np.exp(q + (w[d] + em) - se) for d, q in enumerate((p, p1, p2))]
        c = [G * E[d] for d in range(3)]
        np.add.at(grad[t - 1], lab, (c[0] + c[1] + c[2]) * delta)
        G = c[0] + np.concatenate([c[1][1:], [0.0]]) + np.concatenate([c[2][2:], [0.0, 0.0]])
    return out, grad


def backward_banded(e, w, lab, accept, S, K, delta=1.0):
    """Backward sweep, K levels per barrier.  Lane l of warp q holds node q*own + l - 2: lanes 0, 1
    only supply saved scores to lanes 2, 3; lanes 32-2K .. 31 are the halo of the gradient recursion
    (the dependencies come from u+1, u+2) and go stale two per level."""
    T, C = e.shape
    N = w.shape[1]
    H = 2 * K
    own = 30 - H
    W = -(-N // own)
    lane = np.arange(32)[None, :]
    node = np.arange(W)[:, None] * own + lane - 2
    ok = (node >= 0) & (node < N)
    nc = np.clip(node, 0, N - 1)
    is_own = (lane >= 2) & (lane < 32 - H)
    wl = np.where(ok[None], w[:, nc], NINF)
    labl = lab[nc]
    grad = np.zeros((T, C))
    sa = np.where(accept, S[T], NINF)
    mx = sa.max()
    out = mx + np.log(np.exp(sa - mx).sum())
    Gfull = np.where(accept, np.exp(S[T] - out), 0.0)
    G = np.where(ok, Gfull[nc], 0.0)  # seeds: every lane can compute its own
    row = np.zeros(N)
    t0 = T
    while t0 >= 1:
        for j in range(min(K, t0)):
            t = t0 - j
            s_t = np.where(ok, S[t][nc], NINF)
            s_p = np.where(ok, S[t - 1][nc], NINF)
            sp1, sp2 = shfl_up(s_p, 1), shfl_up(s_p, 2)
            em = e[t - 1][labl]
            se = np.where(np.isinf(s_t), np.inf, s_t)
            E = [np.exp(q + (wl[d] + em) - se) for d, q in enumerate((s_p, sp1, sp2))]
            c = [G * E[d] for d in range(3)]
            m = ok & is_own
            np.add.at(grad[t - 1], labl[m], ((c[0] + c[1] + c[2]) * delta)[m])
            G = c[0] + shfl_down(c[1], 1) + shfl_down(c[2], 2)
        # block boundary: own lanes publish, (barrier), the high halo re-reads
        m = ok & is_own
        row[:] = 0.0
        row[node[m]] = G[m]
        G = np.where(is_own, G, np.where(ok, row[nc], 0.0))
        t0 -= K
    return out, grad


def check(seed, T, C, U, K):
    rng = np.random.default_rng(seed)
    e = rng.uniform(-5, 5, (T, C))
    target = rng.integers(1, C, U)
    if U > 2:
        target[1] = target[0]  # a repeated label: no skip arc there
    w, lab, start, accept = ctc_band(target)
    S0 = forward_plain(e, w, lab, start)
    S1 = forward_banded(e, w, lab, start, K)
    same = (np.isinf(S0) & np.isinf(S1)) | np.isclose(S0, S1, rtol=1e-12, atol=1e-12)
    assert same.all(), ("forward", seed, T, C, U, K, np.argwhere(~same)[:5])
    o0, g0 = backward_plain(e, w, lab, accept, S0)
    o1, g1 = backward_banded(e, w, lab, accept, S0, K)
    assert np.isclose(o0, o1)
    assert np.allclose(g0, g1, rtol=1e-9, atol=1e-12), ("backward", seed, T, C, U, K, np.abs(g0 - g1).max())
    # the gradient of the lattice score wrt the emissions sums to one per frame
    assert np.allclose(g1.sum(1), 1.0, atol=1e-9)


if __name__ == "__main__":
    n = 0
    for K in (1, 2, 4):
        for (T, C, U) in ((1, 4, 1), (3, 5, 1), (9, 5, 4), (20, 8, 9), (37, 6, 18), (64, 28, 30), (50, 64, 100 // 4),
                          (120, 16, 60), (41, 64, 20)):
            if T < U:
                continue
            for seed in range(3):
                check(seed, T, C, U, K)
                n += 1
    print("banded schedule: %d cases agree with the per-frame recursion" % n)
