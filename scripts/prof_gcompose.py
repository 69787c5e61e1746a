"""Where the time of gtnb_compose_graphs goes: per-kernel CUDA-event times (context profiling) and wall time of the
C call, for the operand shapes of scripts/bench_gcompose.py, 1 pair and 32 pairs (shared operands)."""
import ctypes as C
import sys
import time

sys.path.insert(0, '.')
import numpy as np

import gtn_b200.gtn as gtn
from gtn_b200 import capi
from scripts.bench_gcompose import cases
from tests.test_host_api import ours_arrays

L = capi.lib()
ctx = capi.Ctx(0)


def view(g):
    a = ours_arrays(g)
    l = gtn.arc_lists(g)
    return capi.make_view(a["flags"], a["src"], a["dst"], a["ilabel"], a["olabel"], a["w"], l["in_ptr"], l["in_arcs"],
                          l["out_ptr"], l["out_arcs"], l["start"], l["accept"])


for name, (a, b, inter) in cases.items():
    va, vb = view(a), view(b)
    s1 = (a.ilabel_sorted() or a.olabel_sorted()) if inter else a.olabel_sorted()
    s2 = (b.ilabel_sorted() or b.olabel_sorted()) if inter else b.ilabel_sorted()
    kind = 3 if (s1 and s2) else (1 if s1 else (2 if s2 else 0))
    for n in (1, 32):
        kinds = np.full(n, kind, np.int32)
        A1 = (capi.GraphView * 1)(va[0])
        B1 = (capi.GraphView * 1)(vb[0])
        for rep in range(3):
            ctx.profile(True)
            h = capi._vp()
            t0 = time.perf_counter()
            rc = L.gtnb_compose_graphs(ctx.h, n, A1, 1, B1, 1, kinds.ctypes.data_as(capi._i32p), C.byref(h))
            ms = (time.perf_counter() - t0) * 1e3
            ctx._check(rc)
            prof = ctx.profile_read()
            L.gtnb_composed_destroy(ctx.h, h)
        print(name, "pairs", n, "call ms %.3f" % ms, {k: round(v[1], 3) for k, v in prof.items()}, flush=True)
