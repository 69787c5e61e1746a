"""Round-2 first measurement: gtnb_ctc_loss at BASELINE config 2 with the per-frame sweeps of
k_implicit.cu (banded = 0) and the temporally blocked ones of k_banded.cu at K = 1, 2, 4, 8 frames per
barrier.  Device-resident buffers, L2 flushed, CUDA events; per-kernel times from gtnb_profile_*.
Also checks every variant against banded = 0 (losses and gradients) before timing it.

    GTNB_EXPERIMENTAL=1 python -m pytest tests/test_gpu_parity.py -q -k banded   # parity first
    python scripts/bench_banded.py > gpurun_out/banded.json
"""
import json
import sys

sys.path.insert(0, '.')
import numpy as np

import bench
from gtn_b200 import capi

L = capi.lib()
ctx = capi.Ctx(0)
i32p, f32p = capi._i32p, capi._f32p
B, T, C, U = 256, 1000, 64, 100
e, tg = bench.make_inputs(0, B, T, C, U)
lens = np.full(B, U, np.int32)
cat = np.ascontiguousarray(np.concatenate(tg), np.int32)
e_dev = ctx.to_device(e)
g_dev = ctx.alloc(e.nbytes)
losses = np.zeros(B, np.float32)


def step():
    ctx._check(L.gtnb_ctc_loss(ctx.h, B, T, C, e_dev.ptr, 1, None, cat.ctypes.data_as(i32p), lens.ctypes.data_as(i32p),
                               0, losses.ctypes.data_as(f32p), g_dev.ptr, 1))


def grads():
    return g_dev.download((B, T, C))


out = {}
ref = None
for K in (0, 1, 2, 4, 8):
    ctx.set_flag("banded", K)
    for _ in range(3):
        step()
    res = {"losses_head": losses[:3].tolist()}
    g = grads()
    if K == 0:
        ref = (losses.copy(), g)
    else:
        res["max_abs_loss_diff"] = float(np.abs(losses - ref[0]).max())
        res["max_abs_grad_diff"] = float(np.abs(g - ref[1]).max())
    ctx.profile(True)
    ctx.profile_read()
    ts = []
    for _ in range(30):
        ctx.flush_l2()
        ctx.timer_start()
        step()
        ts.append(ctx.timer_stop())
    pr = ctx.profile_read()
    ctx.profile(False)
    res.update({"ms_median": float(np.median(ts)), "ms_min": float(np.min(ts)),
                "kernels_ms": {k: v[1] / v[0] for k, v in pr.items() if k != "flush_l2"}})
    out["banded_%d" % K] = res
    print("banded", K, json.dumps(res), file=sys.stderr, flush=True)
ctx.set_flag("banded", 0)
print(json.dumps(out, indent=1))
