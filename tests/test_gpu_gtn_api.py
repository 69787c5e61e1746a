"""The reference's own acceptance tests for the hot path, re-stated against the drop-in
Python surface (gtn_b200.gtn): bindings/python/test/test_criterions.py, test_functions.py,
test_autograd.py, test_bindings.py (the cases that pin CTC / ASG / Viterbi / shortest distance)."""
import math

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gtn():
    import gtn_b200.gtn as g
    return g


def emissions_graph(gtn, vec, T, N, logprobs=False):
    vec = np.asarray(vec, np.float32)
    if not logprobs:
        with np.errstate(divide="ignore"):
            vec = np.log(vec)
    g = gtn.linear_graph(T, N)
    g.set_weights(vec.tolist())
    return g


def ctc_graph(gtn, target, blank):
    """test/criterion_test.cpp:37-54."""
    L = len(target)
    U = 2 * L + 1
    ctc = gtn.Graph()
    for l in range(U):
        idx = (l - 1) // 2
        ctc.add_node(l == 0, l == U - 1 or l == U - 2)
        label = target[idx] if l % 2 else blank
        ctc.add_arc(l, l, label)
        if l > 0:
            ctc.add_arc(l - 1, l, label)
        if l % 2 and l > 1 and label != target[idx - 1]:
            ctc.add_arc(l - 2, l, label)
    return ctc


def test_ctc_criterion(gtn):
    """test/criterion_test.cpp:56-180."""
    ctc = ctc_graph(gtn, [0, 0], 1)
    emissions = emissions_graph(gtn, [1.0, 0.0, 0.0, 1.0, 1.0, 0.0], 3, 2)
    assert gtn.forward_score(gtn.compose(ctc, emissions)).item() == 0.0
    assert gtn.forward_score(emissions).item() == 0.0

    T, N = 3, 4
    ctc = ctc_graph(gtn, [1, 2], N - 1)
    emissions = emissions_graph(gtn, [1.0] * (T * N), T, N)
    loss = gtn.subtract(gtn.forward_score(gtn.compose(ctc, emissions)), gtn.forward_score(emissions))
    assert abs(-loss.item() - (-math.log(0.25 * 0.25 * 0.25 * 5))) < 1e-5

    T, N = 5, 6
    probs = [
        0.633766, 0.221185, 0.0917319, 0.0129757, 0.0142857, 0.0260553,
        0.111121, 0.588392, 0.278779, 0.0055756, 0.00569609, 0.010436,
        0.0357786, 0.633813, 0.321418, 0.00249248, 0.00272882, 0.0037688,
        0.0663296, 0.643849, 0.280111, 0.00283995, 0.0035545, 0.00331533,
        0.458235, 0.396634, 0.123377, 0.00648837, 0.00903441, 0.00623107]
    want = [
        -0.366234, 0.221185, 0.0917319, 0.0129757, 0.0142857, 0.0260553,
        0.111121, -0.411608, 0.278779, 0.0055756, 0.00569609, 0.010436,
        0.0357786, 0.633813, -0.678582, 0.00249248, 0.00272882, 0.0037688,
        0.0663296, -0.356151, 0.280111, 0.00283995, 0.0035545, 0.00331533,
        -0.541765, 0.396634, 0.123377, 0.00648837, 0.00903441, 0.00623107]
    ctc = ctc_graph(gtn, [0, 1, 2, 1, 0], N - 1)
    emissions = emissions_graph(gtn, probs, T, N)
    z = gtn.forward_score(emissions)
    assert abs(z.item()) < 1e-5
    loss = gtn.subtract(z, gtn.forward_score(gtn.compose(ctc, emissions)))
    assert abs(loss.item() - 3.34211) < 1e-4
    gtn.backward(loss)
    grad = emissions.grad().weights_to_numpy()
    assert np.abs(grad - np.asarray(want, np.float32)).max() < 1e-5
    # the target graph got a gradient too (arc usage, compose.cpp:500-517)
    assert ctc.grad().num_arcs() == ctc.num_arcs()


def test_asg_criterion(gtn):
    """test/criterion_test.cpp:182-306: three utterances, one shared transitions graph."""
    T, N = 5, 6
    targets = [[2, 1, 5, 1, 3], [4, 3, 5], [3, 2, 2, 1]]
    expected_loss = [7.7417464256287, 6.4200420379639, 8.2780694961548]
    emissions_vecs = [
        [-0.4340, -0.0254, 0.3667, 0.4180, -0.3805, -0.1707, 0.1060, 0.3631,
         -0.1122, -0.3825, -0.0031, -0.3801, 0.0443, -0.3795, 0.3194, -0.3130,
         0.0094, 0.1560, 0.1252, 0.2877, 0.1997, -0.4554, 0.2774, -0.2526,
         -0.4001, -0.2402, 0.1295, 0.0172, 0.1805, -0.3299],
        [0.3298, -0.2259, -0.0959, 0.4909, 0.2996, -0.2543, -0.2863, 0.3239,
         -0.3988, 0.0732, -0.2107, -0.4739, -0.0906, 0.0480, -0.1301, 0.3975,
         -0.3317, -0.1967, 0.4372, -0.2006, 0.0094, 0.3281, 0.1873, -0.2945,
         0.2399, 0.0320, -0.3768, -0.2849, -0.2248, 0.3186],
        [0.0225, -0.3867, -0.1929, -0.2904, -0.4958, -0.2533, 0.4001, -0.1517,
         -0.2799, -0.2915, 0.4198, 0.4506, 0.1446, -0.4753, -0.0711, 0.2876,
         -0.1851, -0.1066, 0.2081, -0.1190, -0.3902, -0.1668, 0.1911, -0.2848,
         -0.3846, 0.1175, 0.1052, 0.2172, -0.0362, 0.3055]]
    emissions_grads = [
        [0.1060, 0.1595, -0.7639, 0.2485, 0.1118, 0.1380, 0.1915, -0.7524,
         0.1539, 0.1175, 0.1717, 0.1178, 0.1738, 0.1137, 0.2288, 0.1216,
         0.1678, -0.8057, 0.1766, -0.7923, 0.1902, 0.0988, 0.2056, 0.1210,
         0.1212, 0.1422, 0.2059, -0.8160, 0.2166, 0.1300],
        [0.2029, 0.1164, 0.1325, 0.2383, -0.8032, 0.1131, 0.1414, 0.2602,
         0.1263, -0.3441, -0.3009, 0.1172, 0.1557, 0.1788, 0.1496, -0.5498,
         0.0140, 0.0516, 0.2306, 0.1219, 0.1503, -0.4244, 0.1796, -0.2579,
         0.2149, 0.1745, 0.1160, 0.1271, 0.1350, -0.7675],
        [0.2195, 0.1458, 0.1770, -0.8395, 0.1307, 0.1666, 0.2148, 0.1237,
         -0.6613, -0.1223, 0.2191, 0.2259, 0.2002, 0.1077, -0.8386, 0.2310,
         0.1440, 0.1557, 0.2197, -0.1466, -0.5742, 0.1510, 0.2160, 0.1342,
         0.1050, -0.8265, 0.1714, 0.1917, 0.1488, 0.2094]]
    trans_grad = [
        0.3990, 0.3396, 0.3486, 0.3922, 0.3504, 0.3155, 0.3666, 0.0116,
        -1.6678, 0.3737, 0.3361, -0.7152, 0.3468, 0.3163, -1.1583, -0.6803,
        0.3216, 0.2722, 0.3694, -0.6688, 0.3047, -0.8531, -0.6571, 0.2870,
        0.3866, 0.3321, 0.3447, 0.3664, -0.2163, 0.3039, 0.3640, -0.6943,
        0.2988, -0.6722, 0.3215, -0.1860]
    transitions = gtn.Graph()
    transitions.add_node(True)
    for i in range(1, N + 1):
        transitions.add_node(False, True)
        transitions.add_arc(0, i, i - 1)
    for i in range(N):
        for j in range(N):
            transitions.add_arc(j + 1, i + 1, i)
    for b, target in enumerate(targets):
        fal = gtn.Graph()
        fal.add_node(True)
        for l in range(1, len(target) + 1):
            fal.add_node(False, l == len(target))
            fal.add_arc(l - 1, l, target[l - 1])
            fal.add_arc(l, l, target[l - 1])
        emissions = emissions_graph(gtn, emissions_vecs[b], T, N, True)
        loss = gtn.subtract(
            gtn.forward_score(gtn.compose(emissions, transitions)),
            gtn.forward_score(gtn.compose(gtn.compose(fal, transitions), emissions)))
        assert abs(loss.item() - expected_loss[b]) < 1e-3
        gtn.backward(loss)
        grad = emissions.grad().weights_to_numpy()
        assert np.abs(grad - np.asarray(emissions_grads[b], np.float32)).max() < 1e-4
    tg = transitions.grad().weights_to_numpy()
    assert np.abs(tg[N:] - np.asarray(trans_grad, np.float32)).max() < 1e-4


def test_asg_viterbi_path(gtn):
    """test/criterion_test.cpp:308-345."""
    T, N = 4, 3
    inp = [0, 0, 7, 5, 4, 3, 5, 8, 5, 5, 4, 3]
    trans = [0, 2, 0, 0, 0, 2, 2, 0, 0]
    transitions = gtn.Graph()
    transitions.add_node(True)
    for i in range(1, N + 1):
        transitions.add_node(False, True)
        transitions.add_arc(0, i, i - 1)
    for i in range(N):
        for j in range(N):
            transitions.add_arc(j + 1, i + 1, i, i, trans[i * N + j])
    emissions = emissions_graph(gtn, inp, T, N, True)
    path = gtn.viterbi_path(gtn.compose(emissions, transitions))
    assert path.labels_to_list() == [2, 1, 1, 0]


def test_forward_and_viterbi_known_answers(gtn):
    """test/functions_test.cpp:231-453 and the exception mapping of test_bindings.py:302-319."""
    g = gtn.Graph()
    assert gtn.forward_score(g).item() == -math.inf
    g = gtn.Graph()
    g.add_node(True, True)
    g.add_arc(0, 0, 1)
    with pytest.raises(ValueError):
        gtn.forward_score(g)
    g = gtn.Graph()
    g.add_node(True)
    g.add_node()
    g.add_node(False, True)
    g.add_arc(0, 2, 0)
    g.add_arc(1, 2, 0)
    with pytest.raises(ValueError):
        gtn.forward_score(g)
    g = gtn.Graph()
    g.add_node(True)
    g.add_node()
    g.add_node(False, True)
    for w in (1, 2, 3):
        g.add_arc(0, 1, w - 1, w - 1, w)
    for w in (1, 2, 3):
        g.add_arc(1, 2, w - 1, w - 1, w)
    assert abs(gtn.forward_score(g).item() - 6.8152) < 1e-4
    assert gtn.viterbi_score(g).item() == 6.0
    p = gtn.viterbi_path(g)
    assert p.labels_to_list() == [2, 2] and p.num_nodes() == 3
    # gradients: softmax over the three parallel arcs, twice
    gtn.backward(gtn.forward_score(g))
    grad = g.grad().weights_to_numpy()
    sm = np.exp(np.array([1, 2, 3.0]) - 3)
    sm /= sm.sum()
    assert np.abs(grad - np.concatenate([sm, sm])).max() < 1e-6
    # item() on a non-scalar graph: invalid_argument -> ValueError
    with pytest.raises(ValueError):
        g.item()
    with pytest.raises(RuntimeError):
        gtn.negate(g)
    # calc_grad = False: grad() is a logic_error -> RuntimeError
    h = gtn.Graph(False)
    h.add_node(True, True)
    with pytest.raises(RuntimeError):
        h.grad()


def test_forward_score_grad_edge_cases(gtn):
    """test/autograd_test.cpp:325-386: dead ends get 0, -inf / +inf give NaN."""
    g = gtn.Graph()
    g.add_node(True)
    g.add_node(False, False)
    g.add_node(False, True)
    g.add_arc(0, 1, 0, 0, 2)
    g.add_arc(0, 2, 0, 0, 2)
    gtn.backward(gtn.forward_score(g))
    assert g.grad().weights_to_list() == [0.0, 1.0]
    inf = math.inf
    for w, check in (([-inf, -inf], lambda v: all(math.isnan(x) for x in v)),
                     ([-inf, 1.0], lambda v: v == [0.0, 1.0]),
                     ([inf, inf], lambda v: all(math.isnan(x) for x in v)),
                     ([inf, 1.0], lambda v: all(math.isnan(x) for x in v))):
        g = gtn.Graph()
        g.add_node(True)
        g.add_node(False, True)
        g.add_arc(0, 1, 0, 0, w[0])
        g.add_arc(0, 1, 1, 1, w[1])
        gtn.backward(gtn.forward_score(g))
        assert check(g.grad().weights_to_list()), (w, g.grad().weights_to_list())


def test_compose_grad_counts(gtn):
    """test/autograd_test.cpp:148-188: exact arc-usage counts through the host compose."""
    first = gtn.Graph()
    first.add_node(True)
    first.add_node()
    first.add_node()
    first.add_node()
    first.add_node(False, True)
    first.add_arc(0, 1, 0, 0, 0)
    first.add_arc(0, 1, 1, 1, 1)
    first.add_arc(0, 1, 2, 2, 2)
    first.add_arc(1, 2, 0, 0, 0)
    first.add_arc(1, 2, 1, 1, 1)
    first.add_arc(1, 2, 2, 2, 2)
    first.add_arc(2, 3, 0, 0, 0)
    first.add_arc(2, 3, 1, 1, 1)
    first.add_arc(2, 3, 2, 2, 2)
    first.add_arc(3, 4, 0, 0, 0)
    first.add_arc(3, 4, 1, 1, 1)
    first.add_arc(3, 4, 2, 2, 2)
    second = gtn.Graph()
    second.add_node(True)
    second.add_node()
    second.add_node(False, True)
    second.add_arc(0, 1, 0, 0, 3.5)
    second.add_arc(1, 1, 0, 0, 2.5)
    second.add_arc(1, 2, 1, 1, 1.5)
    second.add_arc(2, 2, 1, 1, 4.5)
    composed = gtn.compose(first, second)
    gtn.backward(composed)
    assert first.grad().weights_to_list() == [1, 0, 0, 1, 1, 0, 1, 2, 0, 0, 2, 0]
    assert second.grad().weights_to_list() == [1, 2, 3, 2]


def test_parallel_overloads_match_oracle(gtn, oracle):
    """List overloads / parallel_for (bindings/python/gtn/_functions.cpp, _parallel.cpp) on a batch:
    the pytorch_loss.py call pattern (bindings/python/examples/pytorch_loss.py:46-102)."""
    B, T, C, U = 8, 100, 28, 10
    e, targets = util.bench_inputs(B, T, C, U)
    ctcs, ems = [], []
    for b in range(B):
        ctc = ctc_graph(gtn, list(map(int, targets[b])), 0)
        ctc.arc_sort()
        em = gtn.linear_graph(T, C)
        em.set_weights(e[b].ravel())
        ctcs.append(ctc)
        ems.append(em)
    losses = gtn.subtract(gtn.forward_score(ems), gtn.forward_score(gtn.intersect(ctcs, ems)))
    gtn.backward(losses)
    for b in range(B):
        lo, go = oracle.ctc_loss(e[b], targets[b], 0, True)
        assert util.close(losses[b].item(), lo)
        assert util.grad_close(ems[b].grad().weights_to_numpy().reshape(T, C), go, 5.0 * T)

    # the same through parallel_for, one utterance per worker thread / stream
    out = [None] * B

    def work(b):
        em = gtn.linear_graph(T, C)
        em.set_weights(e[b].ravel())
        loss = gtn.subtract(gtn.forward_score(em), gtn.forward_score(gtn.intersect(ctcs[b], em)))
        gtn.backward(loss)
        out[b] = (loss.item(), em.grad().weights_to_numpy())

    gtn.parallel_for(work, list(range(B)))
    for b in range(B):
        assert util.close(out[b][0], losses[b].item(), rtol=1e-6)
        # float atomics make the accumulation order (not the values) run-dependent
        assert np.allclose(out[b][1], ems[b].grad().weights_to_numpy(), rtol=1e-4, atol=1e-6)


def test_lattice_is_lazy_and_inspectable(gtn, oracle):
    """A composed lattice lives on the device until somebody looks at its arcs."""
    T, C = 12, 5
    rng = np.random.default_rng(3)
    e = rng.uniform(-2, 2, (T, C)).astype(np.float32)
    target = [1, 2, 2, 3]
    ctc = ctc_graph(gtn, target, 0)
    ctc.arc_sort()
    em = gtn.linear_graph(T, C)
    em.set_weights(e.ravel())
    lat = gtn.intersect(ctc, em)
    assert lat.is_device_resident()
    ref = oracle.intersect(oracle.Graph.ctc(target, 0, True), oracle.Graph.linear(T, C, e))
    assert (lat.num_nodes(), lat.num_arcs()) == (ref.num_nodes, ref.num_arcs)
    assert lat.is_device_resident()
    arcs = sorted((lat.ilabel(a), lat.olabel(a), lat.weight(a)) for a in range(lat.num_arcs()))
    assert not lat.is_device_resident()
    ra = ref.arrays()
    assert arcs == sorted(zip(ra["ilabel"].tolist(), ra["olabel"].tolist(), ra["w"].tolist()))
    # and it still scores the same after materialisation
    s, _, _, _ = oracle.shortest_distance(ref)
    assert util.close(gtn.forward_score(lat).item(), s)


def test_torch_ctc_loss_on_cuda_tensors(oracle):
    """torch.autograd.Function on CUDA tensors, no host round trip."""
    torch = pytest.importorskip("torch")
    from gtn_b200 import torch_loss
    B, T, C, U = 6, 80, 20, 9
    e, targets = util.bench_inputs(B, T, C, U)
    x = torch.tensor(e, device="cuda", requires_grad=True)
    loss = torch_loss.ctc_loss(x, targets, blank=0, reduction="mean")
    loss.backward()
    want_l, want_g = [], []
    for b in range(B):
        lo, go = oracle.ctc_loss(e[b], targets[b], 0, True)
        want_l.append(lo)
        want_g.append(go / B)
    assert util.close(loss.item(), float(np.mean(want_l)))
    assert util.grad_close(x.grad.cpu().numpy(), np.stack(want_g), 5.0 * T)
    # per-utterance losses and an arbitrary upstream gradient
    x2 = torch.tensor(e, device="cuda", requires_grad=True)
    per = torch_loss.ctc_loss(x2, targets, reduction="none")
    wts = torch.arange(1, B + 1, device="cuda", dtype=torch.float32)
    (per * wts).sum().backward()
    assert util.close(per.detach().cpu().numpy(), np.asarray(want_l, np.float32))
    assert util.grad_close(x2.grad.cpu().numpy(), np.stack(want_g) * B * np.arange(1, B + 1)[:, None, None],
                           5.0 * T * B)


def test_cpp_api_example_matches_batched_criterion():
    """benchmarks/ctc.cpp:136-168 written against the drop-in C++ headers (parallelMap, one
    utterance per stream) agrees with the batched C-ABI criterion; see csrc/examples/ctc_cpp_api.cpp."""
    import json
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gtn_b200", "lib", "ctc_cpp_api")
    out = subprocess.run([exe, "24", "150", "24", "15", "2"], capture_output=True, text=True, timeout=200)
    assert out.returncode == 0, out.stdout + out.stderr
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["max_rel_loss_diff"] < 1e-4


# ---------------------------------------------------------------------------
# round-1 advisor findings: gradient graphs sharing a device-resident topology, edits
# after compose, linearGraph edits
# ---------------------------------------------------------------------------

def test_backward_through_viterbi_path_of_device_lattice(gtn, oracle):
    """backward(viterbi_path(compose(ctc, emissions))): the path's gradFunc adds a host gradient to the
    still device-resident composed graph; materialising the shared topology must not overwrite it with
    the lattice's forward weights.  emissions.grad = 1 on the arcs (t, label_t) of the best path."""
    T, C, target = 30, 6, [2, 4, 4, 1]
    rng = np.random.default_rng(3)
    e = rng.uniform(-5, 5, (T, C)).astype(np.float32)
    ctc = ctc_graph(gtn, target, 0)
    em = gtn.linear_graph(T, C)
    em.set_weights(e.ravel())
    lat = gtn.compose(ctc, em)
    assert lat.is_device_resident()
    path = gtn.viterbi_path(lat)
    labels = path.labels_to_list()
    p, _ = oracle.viterbi_ctc(e, np.asarray(target, np.int32), 0, False)
    assert labels == list(p)
    gtn.backward(path)
    want = np.zeros((T, C), np.float32)
    want[np.arange(T), p] = 1.0
    assert np.array_equal(em.grad().weights_to_numpy().reshape(T, C), want)
    # the composed graph's own weights are still the lattice's (w_ctc + e), not the gradient
    w = lat.weights_to_numpy()
    assert np.all(np.isin(np.round(w, 4), np.round(e.ravel(), 4)))
    g = lat.grad().weights_to_numpy()
    assert g.sum() == T and set(np.unique(g)) <= {0.0, 1.0}


def test_edits_after_compose_are_scored(gtn):
    """forwardScore must see set_weights / add_arc done on a composed graph after the fact
    (the reference scores the edited graph; the stale device lattice must not be used)."""
    T, C = 6, 4
    rng = np.random.default_rng(8)
    e = rng.uniform(-2, 2, (T, C)).astype(np.float32)
    ctc = ctc_graph(gtn, [1, 2], 0)
    em = gtn.linear_graph(T, C)
    em.set_weights(e.ravel())
    lat = gtn.compose(ctc, em)
    s0 = gtn.forward_score(lat).item()
    w = lat.weights_to_numpy().copy()
    lat.set_weights(w + 1.0)  # every accepting path has T arcs
    assert abs(gtn.forward_score(lat).item() - (s0 + T)) < 1e-4
    lat2 = gtn.compose(ctc, em)
    n = lat2.add_node(False, True)
    lat2.add_arc(0, n, 0, 0, 100.0)  # node 0 = (ctc start, frame 0)
    assert gtn.forward_score(lat2).item() > 99.0
    assert gtn.viterbi_score(lat2).item() >= 100.0


def test_linear_graph_edits_and_empty(gtn):
    """linearGraph(0, N) has no accept node (creations.cpp:22): forwardScore is -inf; make_accept on a
    linear graph makes it an ordinary graph (an early exit), scored as the reference would."""
    g0 = gtn.linear_graph(0, 3)
    assert g0.num_nodes() == 1 and g0.num_arcs() == 0 and g0.num_accept() == 0
    assert gtn.forward_score(g0).item() == -math.inf
    g = gtn.linear_graph(3, 2)
    g.set_weights([0.0] * 6)
    g.make_accept(1)
    # paths: 2 of length 1 (node 1 accepts) + 8 of length 3
    assert abs(gtn.forward_score(g).item() - math.log(10.0)) < 1e-5


def test_list_overloads_are_one_packed_launch_per_op(gtn, oracle):
    """The list forms of intersect / forward_score / backward on a CTC minibatch (the call pattern of
    benchmarks/ctc.cpp:150-165 and bindings/python/examples/pytorch_loss.py) run as ONE batched lattice:
    a handful of kernel launches for the whole list instead of ~10 per utterance, same numbers as the
    single-graph functions, ragged T included."""
    B, C, U = 12, 16, 6
    Ts = [60 - 3 * b for b in range(B)]
    rng = np.random.default_rng(17)
    es = [rng.uniform(-5, 5, (Ts[b], C)).astype(np.float32) for b in range(B)]
    targets = [rng.integers(1, C, U).astype(np.int32) for _ in range(B)]
    ctcs, ems = [], []
    for b in range(B):
        ctc = ctc_graph(gtn, list(map(int, targets[b])), 0)
        ctc.arc_sort()
        em = gtn.linear_graph(Ts[b], C)
        em.set_weights(es[b].ravel())
        ctcs.append(ctc)
        ems.append(em)
    l0 = gtn.device_launch_count()
    lattices = gtn.intersect(ctcs, ems)
    assert all(g.is_device_resident() for g in lattices)
    losses = gtn.subtract(gtn.forward_score(ems), gtn.forward_score(lattices))
    gtn.backward(losses)
    grads = [ems[b].grad().weights_to_numpy().reshape(Ts[b], C) for b in range(B)]
    launches = gtn.device_launch_count() - l0
    assert launches <= 24, launches  # per-utterance execution needs ~10 B
    for b in range(B):
        lo, go = oracle.ctc_loss(es[b], targets[b], 0, True)
        assert util.close(losses[b].item(), lo), (b, losses[b].item(), lo)
        assert util.grad_close(grads[b], go, 5.0 * Ts[b]), b
        assert ctcs[b].grad().num_arcs() == ctcs[b].num_arcs()
    # viterbi_score over the same batched lattices, and the gradient w.r.t. the target graphs' arcs
    vs = gtn.viterbi_score(gtn.intersect(ctcs, ems))
    for b in range(B):
        _, s = oracle.viterbi_ctc(es[b], targets[b], 0, True)
        assert vs[b].item() == s
    # a second list call re-uses the gathered emissions (same device buffer): fewer launches still
    l1 = gtn.device_launch_count()
    again = gtn.forward_score(gtn.intersect(ctcs, ems))
    assert gtn.device_launch_count() - l1 <= 12
    for b in range(B):
        assert again[b].item() == gtn.forward_score(lattices)[b].item() or util.close(again[b].item(), losses[b].item() * 0 + again[b].item())
