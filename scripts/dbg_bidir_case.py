import sys
sys.path.insert(0, '.')
import numpy as np
from gtn_b200 import capi
from tests import util
ctx = capi.Ctx(0)
ctx.set_flag("bidir", 1)
for (B, T, C, U, wg) in [(5, 1, 8, 0, True), (4, 120, 16, 9, False), (3, 37, 8, 1, True), (2, 9, 4, 4, True)]:
    e, targets = util.bench_inputs(B, T, C, U, seed=555)
    lens = np.array([T - (5 * b) % max(T // 2, 1) for b in range(B)], np.int32)
    for rep in range(20):
        l, g = ctx.ctc_loss(e, targets, input_lens=lens, want_grad=wg)
    print((B, T, C, U, wg), l[:3])
