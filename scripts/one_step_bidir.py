"""A few gtnb_ctc_loss calls at BASELINE config 2 through k_bidir.cu (for ncu captures)."""
import sys
sys.path.insert(0, '.')
import numpy as np
from gtn_b200 import capi
import bench
L = capi.lib(); ctx = capi.Ctx(0)
ctx.set_flag("bidir", 1)
i32p, f32p = capi._i32p, capi._f32p
B, T, C, U = 256, 1000, 64, 100
e, tg = bench.make_inputs(0, B, T, C, U)
lens = np.full(B, U, np.int32); cat = np.ascontiguousarray(np.concatenate(tg), np.int32)
e_dev = ctx.to_device(e); g_dev = ctx.alloc(e.nbytes); losses = np.zeros(B, np.float32)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    ctx._check(L.gtnb_ctc_loss(ctx.h, B, T, C, e_dev.ptr, 1, None, cat.ctypes.data_as(i32p), lens.ctypes.data_as(i32p), 0, losses.ctypes.data_as(f32p), g_dev.ptr, 1))
print(losses[:3])
