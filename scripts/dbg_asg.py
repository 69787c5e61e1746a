import sys, faulthandler
sys.path.insert(0, '.')
faulthandler.dump_traceback_later(40, exit=True)
import numpy as np
import gtn_b200.gtn as gtn
def P(*a): print(*a, flush=True)
T, N = 5, 6
rng = np.random.default_rng(0)
e = rng.uniform(-1, 1, (T, N)).astype(np.float32)
transitions = gtn.Graph()
transitions.add_node(True)
for i in range(1, N + 1):
    transitions.add_node(False, True); transitions.add_arc(0, i, i - 1)
for i in range(N):
    for j in range(N):
        transitions.add_arc(j + 1, i + 1, i)
target = [2, 1, 5, 1, 3]
fal = gtn.Graph(); fal.add_node(True)
for l in range(1, len(target) + 1):
    fal.add_node(False, l == len(target)); fal.add_arc(l - 1, l, target[l - 1]); fal.add_arc(l, l, target[l - 1])
em = gtn.linear_graph(T, N); em.set_weights(e.ravel())
P("compose(em, trans)")
den = gtn.compose(em, transitions)
P("sizes", den.num_nodes(), den.num_arcs())
fd = gtn.forward_score(den); P("fwd den", fd.item())
P("compose(fal, trans)")
ft = gtn.compose(fal, transitions); P("ft", ft.num_nodes(), ft.num_arcs())
num = gtn.compose(ft, em); P("num", num.num_nodes(), num.num_arcs())
fn = gtn.forward_score(num); P("fwd num", fn.item())
loss = gtn.subtract(fd, fn); P("loss", loss.item())
gtn.backward(loss); P("backward done")
P(em.grad().weights_to_numpy()[:6]); P(transitions.grad().weights_to_numpy()[:8])
