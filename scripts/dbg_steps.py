import sys, time
sys.path.insert(0, '.')
import numpy as np, ctypes as Ct
from gtn_b200 import capi
import bench
L = capi.lib(); ctx = capi.Ctx(0)
B,T,C,U = 256,1000,64,100
e,tg = bench.make_inputs(0,B,T,C,U)
lens = np.full(B,U,np.int32); cat = np.ascontiguousarray(np.concatenate(tg),np.int32)
e_dev = ctx.to_device(e); g_dev = ctx.alloc(e.nbytes); losses = np.zeros(B,np.float32)
i32p,f32p = capi._i32p, capi._f32p
def step():
    ctx._check(L.gtnb_ctc_loss(ctx.h,B,T,C,e_dev.ptr,1,None,cat.ctypes.data_as(i32p),lens.ctypes.data_as(i32p),0,losses.ctypes.data_as(f32p),g_dev.ptr,1))
for flush in (False, True, True):
    ts=[]; ws=[]
    for i in range(60):
        if flush: ctx.flush_l2()
        w0=time.perf_counter(); ctx.timer_start(); step(); ts.append(ctx.timer_stop()); ws.append((time.perf_counter()-w0)*1e3)
    ts=np.array(ts); print('flush',flush,'mean %.2f median %.2f min %.2f max %.2f'%(ts.mean(),np.median(ts),ts.min(),ts.max()), 'n>4ms', int((ts>4).sum()), [round(x,1) for x in ts[:25]])
