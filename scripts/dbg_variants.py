import sys, time, subprocess, os
sys.path.insert(0, '.')
variant = sys.argv[1] if len(sys.argv) > 1 else None
if variant is None:
    for v in ["base", "torch", "nvml", "prof", "pinned", "nvmlsample"]:
        out = subprocess.run([sys.executable, __file__, v], capture_output=True, text=True).stdout.strip().splitlines()
        print(v, out[-1] if out else "no output", flush=True)
    sys.exit(0)
import numpy as np, ctypes as Ct
if variant == "torch":
    import torch
    torch.cuda.set_device(0)
from gtn_b200 import capi
import bench
L = capi.lib(); ctx = capi.Ctx(0)
B,T,C,U = 256,1000,64,100
e,tg = bench.make_inputs(0,B,T,C,U)
lens = np.full(B,U,np.int32); cat = np.ascontiguousarray(np.concatenate(tg),np.int32)
e_dev = ctx.to_device(e); g_dev = ctx.alloc(e.nbytes); losses = np.zeros(B,np.float32)
i32p,f32p = capi._i32p, capi._f32p
sampler = None
if variant in ("nvml", "nvmlsample"):
    sampler = bench.ClockSampler(0)
if variant == "pinned":
    a, b = Ct.c_void_p(), Ct.c_void_p()
    L.gtnb_host_alloc(ctx.h, e.nbytes, Ct.byref(a)); L.gtnb_host_alloc(ctx.h, e.nbytes, Ct.byref(b))
if variant == "prof":
    ctx.profile(True)
def step():
    ctx._check(L.gtnb_ctc_loss(ctx.h,B,T,C,e_dev.ptr,1,None,cat.ctypes.data_as(i32p),lens.ctypes.data_as(i32p),0,losses.ctypes.data_as(f32p),g_dev.ptr,1))
for _ in range(3): step()
ts=[]
for i in range(100):
    ctx.flush_l2(); ctx.timer_start(); step(); ts.append(ctx.timer_stop())
    if variant == "nvmlsample" and i % 12 == 6: sampler.sample()
ts=np.array(ts); print('mean %.2f median %.2f min %.2f max %.2f n>4ms %d'%(ts.mean(),np.median(ts),ts.min(),ts.max(), int((ts>4).sum())))
