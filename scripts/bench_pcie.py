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
# both directions at once (two streams): what full duplex gives on this host
import torch
a = torch.empty(n // 4, dtype=torch.float32).pin_memory(); b_ = torch.empty(n // 4, dtype=torch.float32).pin_memory()
da = torch.empty(n // 4, dtype=torch.float32, device="cuda"); db = torch.empty(n // 4, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for rep in range(4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    with torch.cuda.stream(s1):
        da.copy_(a, non_blocking=True)
    with torch.cuda.stream(s2):
        b_.copy_(db, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    print("duplex 65.5 MB each way: %.3f ms" % e0.elapsed_time(e1))
