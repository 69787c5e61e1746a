"""PCIe copy rates through the C ABI (pinned host buffers), and the e2e step split."""
import sys, ctypes as Ct
sys.path.insert(0, '.')
import numpy as np
from gtn_b200 import capi
import bench
L = capi.lib(); ctx = capi.Ctx(0)
n = 65536000
hp = Ct.c_void_p(); ctx._check(L.gtnb_host_alloc(ctx.h, n, Ct.byref(hp)))
hq = Ct.c_void_p(); ctx._check(L.gtnb_host_alloc(ctx.h, n, Ct.byref(hq)))
Ct.memset(hp.value, 1, n)
d = ctx.alloc(n)
for name, fn in (("h2d", lambda: L.gtnb_memcpy_h2d(ctx.h, d.ptr, hp, n)), ("d2h", lambda: L.gtnb_memcpy_d2h(ctx.h, hq, d.ptr, n))):
    ts = []
    for _ in range(6):
        ctx.timer_start(); ctx._check(fn()); ts.append(ctx.timer_stop())
    print(name, "ms", np.round(ts, 3), "GB/s %.1f" % (n / min(ts) / 1e6))
# e2e with different host-side buffers
i32p, f32p = capi._i32p, capi._f32p
B, T, C, U = 256, 1000, 64, 100
e, tg = bench.make_inputs(0, B, T, C, U)
lens = np.full(B, U, np.int32); cat = np.ascontiguousarray(np.concatenate(tg), np.int32)
Ct.memmove(hp.value, e.ctypes.data, e.nbytes)
losses = np.zeros(B, np.float32)
g_dev = ctx.alloc(e.nbytes); e_dev = ctx.to_device(e)
def run(e_ptr, e_on, g_ptr, g_on, label):
    ts = []
    for _ in range(8):
        ctx.flush_l2(); ctx.timer_start()
        ctx._check(L.gtnb_ctc_loss(ctx.h, B, T, C, e_ptr, e_on, None, cat.ctypes.data_as(i32p), lens.ctypes.data_as(i32p), 0, losses.ctypes.data_as(f32p), g_ptr, g_on))
        ts.append(ctx.timer_stop())
    print(label, np.round(ts, 3))
run(e_dev.ptr, 1, g_dev.ptr, 1, "dev->dev  ")
run(hp.value, 0, g_dev.ptr, 1, "host->dev ")
run(e_dev.ptr, 1, hq.value, 0, "dev->host ")
run(hp.value, 0, hq.value, 0, "host->host")
run(hp.value, 0, None, 0, "host, no grad")
