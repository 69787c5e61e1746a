"""Per-kernel time vs (T, U, B): separates the steady levels from the cone at both lattice ends."""
import sys
sys.path.insert(0, '.')
import numpy as np
from gtn_b200 import capi
import bench
L = capi.lib(); ctx = capi.Ctx(0)
i32p, f32p = capi._i32p, capi._f32p
C = 64
import os
for (B, T, U) in [(256, 1000, 100), (148, 1000, 100), (256, 1000, 10), (592, 1000, 100), (2048, 1000, 100)]:
    e, tg = bench.make_inputs(0, B, T, C, U)
    lens = np.full(B, U, np.int32); cat = np.ascontiguousarray(np.concatenate(tg), np.int32)
    e_dev = ctx.to_device(e); g_dev = ctx.alloc(e.nbytes); losses = np.zeros(B, np.float32)
    def step():
        ctx._check(L.gtnb_ctc_loss(ctx.h, B, T, C, e_dev.ptr, 1, None, cat.ctypes.data_as(i32p), lens.ctypes.data_as(i32p), 0, losses.ctypes.data_as(f32p), g_dev.ptr, 1))
    for _ in range(3): step()
    ctx.profile(True); ctx.profile_read()
    n = 10
    for _ in range(n): ctx.flush_l2(); step()
    pr = ctx.profile_read(); ctx.profile(False)
    print(B, T, U, ' '.join('%s=%.3f' % (k.replace('compose_', 'c_').replace('sd_', ''), v[1] / v[0]) for k, v in pr.items() if k != 'flush_l2'), flush=True)
    del e_dev, g_dev
