"""Where the drop-in list API spends its time at BASELINE config 2 (phase by phase, wall clock)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
import gtn_b200.gtn as gtn
B, T, C, U = 256, 1000, 64, 100
e, tg = bench.make_inputs(0, B, T, C, U)
x = torch.tensor(e, device="cuda")
ctcs = []
for t in tg:
    L = 2 * len(t) + 1
    g = gtn.Graph(False)
    for l in range(L):
        g.add_node(l == 0, l == L - 1 or l == L - 2)
        label = int(t[(l - 1) // 2]) if l % 2 else 0
        g.add_arc(l, l, label)
        if l > 0: g.add_arc(l - 1, l, label)
        if l % 2 and l > 1 and label != int(t[(l - 3) // 2]): g.add_arc(l - 2, l, label)
    g.arc_sort(); ctcs.append(g)
def tick(name, t0):
    torch.cuda.synchronize(); t1 = time.perf_counter(); acc[name] = acc.get(name, 0) + (t1 - t0); return t1
for rep in range(4):
    acc = {}
    torch.cuda.synchronize(); t = time.perf_counter()
    ems = [gtn.linear_graph(T, C) for _ in range(B)]; t = tick("linear_graph", t)
    for b in range(B): ems[b].set_weights(x[b].data_ptr())
    t = tick("set_weights", t)
    lat = gtn.intersect(ctcs, ems); t = tick("intersect", t)
    z = gtn.forward_score(ems); t = tick("forward_score(ems)", t)
    s = gtn.forward_score(lat); t = tick("forward_score(lat)", t)
    losses = gtn.subtract(z, s); t = tick("subtract", t)
    gtn.backward(losses); t = tick("backward", t)
    g0 = ems[0].grad().weights_to_numpy(); t = tick("first grad (flush)", t)
    gs = [ems[b].grad().weights_to_numpy() for b in range(1, B)]; t = tick("other grads", t)
    vals = [l.item() for l in losses]; t = tick("items", t)
    print({k: round(1e3 * v, 2) for k, v in acc.items()}, "total ms", round(1e3 * sum(acc.values()), 1), flush=True)
