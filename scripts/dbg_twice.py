import sys
sys.path.insert(0, '.')
import numpy as np
from gtn_b200 import capi
from tests import util
ctx = capi.Ctx(0)
for (B,T,C,U) in [(5,60,12,7),(2,5,6,3),(4,100,28,10),(2,5,6,3)]:
    e,tg = util.bench_inputs(B,T,C,U)
    try:
        l,g = ctx.ctc_loss(e,tg)
        print('ok',B,T,C,U,l[:2], flush=True)
    except Exception as ex:
        print('FAIL',B,T,C,U,ex, flush=True)
