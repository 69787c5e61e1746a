"""Regenerate tests/golden/*.npz from the REAL reference (oracle/_ref/libgtn_ref.so).

Run in the build container only (it needs /root/reference to have been compiled
by `make -C oracle ref`):   python tests/golden/make_golden.py
The fixtures are small and committed; tests never read /root/reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from tests import util  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def ctc_f64(e, tg, blank=0):
    """float64 CTC loss / gradient (alpha-beta over the same lattice), the 'truth' that
    both fp32 implementations are measured against."""
    e = e.astype(np.float64)
    T, C = e.shape
    lab = [blank]
    for t in tg:
        lab += [int(t), blank]
    L = len(lab)
    NEG = -np.inf

    def preds(l):
        p = [l]
        if l > 0:
            p.append(l - 1)
        if l % 2 == 1 and l > 1 and lab[l] != lab[l - 2]:
            p.append(l - 2)
        return p

    a = np.full(L, NEG)
    a[0] = 0.0
    A = [a]
    for t in range(T):
        n = np.full(L, NEG)
        for l in range(L):
            c = [a[p] for p in preds(l)]
            m = max(c)
            if m > NEG:
                n[l] = m + np.log(sum(np.exp(x - m) for x in c)) + e[t, lab[l]]
        a = n
        A.append(a)
    fin = [a[L - 1]] + ([a[L - 2]] if L > 1 else [])
    m = max(fin)
    s = m + np.log(sum(np.exp(x - m) for x in fin))
    b = np.full(L, NEG)
    b[L - 1] = 0.0
    if L > 1:
        b[L - 2] = 0.0
    occ = np.zeros((T, C))
    for t in range(T - 1, -1, -1):
        nb = np.full(L, NEG)
        al = A[t]
        for l in range(L):
            if b[l] == NEG:
                continue
            v = e[t, lab[l]] + b[l]
            for lp in preds(l):
                if al[lp] > NEG:
                    occ[t, lab[l]] += np.exp(al[lp] + v - s)
                nb[lp] = np.logaddexp(nb[lp], v)
        b = nb
    z = 0.0
    sm = np.zeros((T, C))
    for t in range(T):
        mx = e[t].max()
        ls = mx + np.log(np.exp(e[t] - mx).sum())
        z += ls
        sm[t] = np.exp(e[t] - ls)
    return z - s, sm - occ


def main():
    assert po.have_ref(), "build oracle/_ref first: make -C oracle ref"
    out = {}
    # BASELINE.json configs[0]
    e, tg = util.bench_inputs(1, 100, 28, 10)
    l, g, _ = po.ref_ctc_batch(e, tg)
    out["c1_loss"], out["c1_grad"] = l, g
    # one utterance of configs[1] + float64 truth
    e, tg = util.bench_inputs(1, 1000, 64, 100)
    l, g, _ = po.ref_ctc_batch(e, tg)
    l64, g64 = ctc_f64(e[0], tg[0])
    out["c2_loss"], out["c2_grad"] = l, g.astype(np.float32)
    out["c2_loss_f64"], out["c2_grad_f64"] = np.float64(l64), g64
    # ASG, 3 utterances sharing transitions (configs[2] shape, small)
    rng = np.random.default_rng(42)
    B, T, C, U = 3, 20, 8, 4
    ea = rng.uniform(-5, 5, (B, T, C)).astype(np.float32)
    tw = rng.uniform(-5, 5, C + C * C).astype(np.float32)
    ta = [rng.integers(0, C, U).astype(np.int32) for _ in range(B)]
    la, ga, tga, _ = po.ref_asg_batch(ea, tw, ta)
    out.update(asg_e=ea, asg_tw=tw, asg_targets=np.stack(ta), asg_loss=la, asg_grad=ga, asg_tgrad=tga)
    # Viterbi over the dense trellis with integer scores: ties everywhere (configs[3] shape, small)
    ev = rng.integers(-3, 4, (4, 30, 6)).astype(np.float32)
    tv = rng.integers(-2, 3, 6 + 36).astype(np.float32)
    pv, sv, _ = po.ref_viterbi_dense_batch(ev, tv)
    out.update(vit_e=ev, vit_tw=tv, vit_paths=pv, vit_scores=sv)
    # CTC forced alignment with integer scores
    tc = [rng.integers(1, 6, 5).astype(np.int32) for _ in range(4)]
    pc, sc, _ = po.ref_viterbi_ctc_batch(ev, tc)
    out.update(fa_targets=np.stack(tc), fa_paths=pc, fa_scores=sc)
    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    err = np.abs(out["c2_grad"][0] - g64)
    print("reference fp32 vs float64 at T=1000: max abs err %.3e, row-sum drift %.3e"
          % (err.max(), np.abs(out["c2_grad"][0].sum(1)).max()))


if __name__ == "__main__":
    main()
