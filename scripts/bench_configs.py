"""Device-resident timings of the other BASELINE.json configs (CUDA events, L2 flushed):
C3 ASG B=128 T=500 C=64 U=50; C4 dense Viterbi B=512 T=2000 C=128; forced alignment
(viterbiPath(intersect(ctc, e))) B=64 T=2000 C=128 U=200; C5 CTC B=2048 on one GPU."""
import sys, json, ctypes as Ct
sys.path.insert(0, '.')
import numpy as np
from gtn_b200 import capi
from tests import util
import bench
L = capi.lib(); ctx = capi.Ctx(0)
i32p, f32p = capi._i32p, capi._f32p
out = {}

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        ctx.flush_l2(); ctx.timer_start(); fn(); ts.append(ctx.timer_stop())
    return float(np.median(ts)), float(np.min(ts))

# ---- C3: ASG
B, T, C, U = 128, 500, 64, 50
e, tg = bench.make_inputs(0, B, T, C, U)
rng = np.random.default_rng(7); tw = rng.uniform(-5, 5, C + C * C).astype(np.float32)
lens = np.full(B, U, np.int32); cat = np.ascontiguousarray(np.concatenate(tg), np.int32)
e_dev = ctx.to_device(e); g_dev = ctx.alloc(e.nbytes); losses = np.zeros(B, np.float32); tgr = np.zeros(C + C * C, np.float32)
def asg():
    ctx._check(L.gtnb_asg_loss(ctx.h, B, T, C, e_dev.ptr, 1, tw.ctypes.data_as(f32p), cat.ctypes.data_as(i32p),
                               lens.ctypes.data_as(i32p), losses.ctypes.data_as(f32p), g_dev.ptr, 1, tgr.ctypes.data_as(f32p)))
for imp in (1, 0):
    ctx.set_flag("implicit", imp)
    ctx.profile(True); ctx.profile_read()
    med, mn = timeit(asg)
    pr = ctx.profile_read(); ctx.profile(False)
    key = "C3_asg_implicit" if imp else "C3_asg_materialised"
    out[key] = {"B": B, "T": T, "C": C, "U": U, "ms_median": med, "ms_min": mn, "utt_per_s": B / (med * 1e-3),
                "kernels_ms": {k: v[1] / v[0] for k, v in pr.items() if k != "flush_l2"}}
    print(key, json.dumps(out[key]), flush=True)
ctx.set_flag("implicit", 1)
e_dev.free(); g_dev.free()

# ---- C4: dense Viterbi
B, T, C = 512, 2000, 128
rng = np.random.default_rng(3)
e = rng.uniform(-5, 5, (B, T, C)).astype(np.float32)
tw = rng.uniform(-5, 5, C + C * C).astype(np.float32)
e_dev = ctx.to_device(e); paths = np.zeros((B, T), np.int32); sc = np.zeros(B, np.float32)
def vit():
    ctx._check(L.gtnb_viterbi_dense(ctx.h, B, T, C, e_dev.ptr, 1, None, tw.ctypes.data_as(f32p), paths.ctypes.data_as(i32p), sc.ctypes.data_as(f32p)))
med, mn = timeit(vit)
out["C4_viterbi_dense"] = {"B": B, "T": T, "C": C, "ms_median": med, "ms_min": mn, "utt_per_s": B / (med * 1e-3)}
print(json.dumps(out["C4_viterbi_dense"]), flush=True)
e_dev.free()

# ---- forced alignment through the lattice API
from oracle import pyoracle as po
B, T, C, U = 64, 2000, 128, 200
e, tg = util.bench_inputs(B, T, C, U, seed=777)
e_dev = ctx.to_device(e)
views = [util.view_of(po.Graph.ctc(t, 0, True)) for t in tg]
def fal():
    lat = ctx.compose_linear(views, [T] * B, C, e_dev, T * C)
    lat.viterbi_path(T)
    lat.free()
med, mn = timeit(fal, n=3, warm=1)
out["forced_alignment"] = {"B": B, "T": T, "C": C, "U": U, "ms_median": med, "ms_min": mn, "utt_per_s": B / (med * 1e-3)}
print(json.dumps(out["forced_alignment"]), flush=True)
e_dev.free()

# ---- C5 on one GPU: CTC B=2048
B, T, C, U = 2048, 1000, 64, 100
e, tg = bench.make_inputs(0, B, T, C, U)
lens = np.full(B, U, np.int32); cat = np.ascontiguousarray(np.concatenate(tg), np.int32)
e_dev = ctx.to_device(e); g_dev = ctx.alloc(e.nbytes); losses = np.zeros(B, np.float32)
def ctc():
    ctx._check(L.gtnb_ctc_loss(ctx.h, B, T, C, e_dev.ptr, 1, None, cat.ctypes.data_as(i32p), lens.ctypes.data_as(i32p), 0, losses.ctypes.data_as(f32p), g_dev.ptr, 1))
ctx.profile(True); ctx.profile_read()
med, mn = timeit(ctc)
pr = ctx.profile_read(); ctx.profile(False)
out["C5_ctc_B2048_one_gpu"] = {"B": B, "ms_median": med, "ms_min": mn, "utt_per_s": B / (med * 1e-3),
                               "kernels_ms": {k: v[1] / v[0] for k, v in pr.items() if k != "flush_l2"}}
print(json.dumps(out["C5_ctc_B2048_one_gpu"]), flush=True)
json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
