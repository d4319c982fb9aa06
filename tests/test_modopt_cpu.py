"""CPU suite of the modularity optimiser.

  * oracle/_ref (the reference's own ModularityOptimizer.cpp, compiled where it lies) reproduces the
    known answers of the reference's tests -- the pin;
  * the committed golden vectors were made by it (tests/golden/make_modopt_golden.py);
  * the step functions the CUDA library launches (seurat_b200/csrc/modopt_core.inl), run serially by
    tests/emul, give exactly those vectors -- labels AND the bits of the modularity -- for every
    window size, including windows of one node (the reference step) and whole-sweep windows.
No GPU is used here; the product library itself is exercised by tests/test_gpu_modopt.py.
"""
import json
import os

import numpy as np
import pytest

import emul_modopt
import modopt_cases as mc
import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "modopt_fixtures.json")) as f:
    GOLD = json.load(f)["cases"]
GRAPHS = {name: build for name, build, _ in mc.small_cases()}
needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")


@needs_ref
def test_ref_build_reproduces_the_reference_tests_known_answers():
    p, i, x = mc.slots(mc.karate())
    for params, want in mc.KARATE_KAT + mc.KARATE_KAT_SLM:
        got, nc, _ = oracle.modularity_cluster_ref(p, i, x, *params)
        assert got.tolist() == want and nc == max(want) + 1


def test_emulated_device_steps_reproduce_the_reference_tests_known_answers():
    p, i, x = mc.slots(mc.karate())
    for params, want in mc.KARATE_KAT + mc.KARATE_KAT_SLM:
        for w0 in (1, 2, 5, 34):
            got, nc, _, _ = emul_modopt.cluster(p, i, x, *params, win0=w0, winmax=34)
            assert got.tolist() == want and nc == max(want) + 1


@pytest.mark.parametrize("case", [c for c in GOLD if not c["graph"].startswith("snn:")],
                         ids=lambda c: c["graph"] + "-" + "-".join(map(str, c["params"])))
def test_emulated_device_steps_match_golden(case):
    p, i, x = mc.slots(GRAPHS[case["graph"]]())
    n = p.size - 1
    for w0, wm in ((1, 1), (3, 64), (256, n), (n, n)):
        lab, nc, q, st = emul_modopt.cluster(p, i, x, *case["params"], win0=w0, winmax=wm)
        assert nc == case["n_clusters"], (w0, wm, st)
        assert mc.digest(lab) == case["sha256"], (w0, wm, st)
        assert q == float.fromhex(case["modularity"]), "modularity must agree to the last bit"
        if w0 == 1 and wm == 1:
            assert st["windows"] == 0, "window size 1 must take the single-node step only"


def test_emulated_device_steps_match_golden_snn_20k():
    from seurat_b200.synthetic import gaussian_mixture
    case = [c for c in GOLD if c["graph"] == "snn:20000:50:7:12:20"]
    X, _ = gaussian_mixture(20000, 50, seed=7, n_centres=12)
    p, i, x = oracle.find_neighbors(X, k=20, prune=1 / 15, method="kdtree")["snn"]
    for c in [case[0], case[1], case[3]]:       # Louvain, refinement, SLM (subnetworks above the one-thread size)
        lab, nc, q, st = emul_modopt.cluster(p, i, x, *c["params"])
        assert nc == c["n_clusters"] and mc.digest(lab) == c["sha256"] and q == float.fromhex(c["modularity"])
        assert st["rounds"] < 12 * max(1, st["windows"]), "the fixed point should need few rounds per window"


@needs_ref
@pytest.mark.parametrize("seed", range(6))
def test_emulated_device_steps_match_the_reference_on_random_graphs(seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(40, 900))
    M = mc.random_graph(n, float(rng.uniform(2, 12)), seed=seed + 50, weighted=bool(seed % 2),
                        symmetric=bool(seed % 3), isolated=int(rng.integers(0, 4)))
    p, i, x = mc.slots(M)
    for _ in range(4):
        mf = int(rng.integers(1, 3))
        res = float(rng.choice([0.01, 0.2, 0.8, 1.0])) if mf == 2 else float(rng.choice([0.1, 0.8, 1.0, 4.0, 40.0]))
        params = (mf, res, int(rng.integers(1, 4)), int(rng.integers(1, 4)), int(rng.integers(1, 6)),
                  int(rng.integers(0, 2 ** 40)))
        want, wc, wq = oracle.modularity_cluster_ref(p, i, x, *params)
        got, gc, gq, st = emul_modopt.cluster(p, i, x, *params, win0=int(rng.integers(1, 64)), winmax=n)
        assert gc == wc and (got == want).all() and gq == wq, (params, st)


def test_python_mirror_argument_checks():
    import seurat_b200 as sb
    with pytest.raises(ValueError, match="Please provide an SNN graph"):
        sb.FindClusters(None)
    with pytest.raises(ValueError, match="algorithm not recognised"):
        sb.FindClusters((np.zeros(2, np.int32), np.zeros(0, np.int32), np.zeros(0)), algorithm=7)
    ids = sb.GroupSingletons([0, 0, 1, 2, 2], (np.array([0, 2, 4, 7, 9, 11], np.int32),
                                               np.array([0, 1, 0, 1, 2, 3, 4, 2, 3, 2, 4], np.int32),
                                               np.array([1, .5, .5, 1, 1, .2, .6, .2, 1, .6, 1])), verbose=False)
    assert ids == [0, 0, 2, 2, 2]
    assert sb.GroupSingletons([0, 0, 1], None, group_singletons=False) == [0, 0, "singleton"]


def test_order_fixed_running_sum_is_exact():
    """seq_sum (binade-wise integer prefix sums) == the element-by-element running sum, bit for bit:
    uniform terms, Jaccard-like rationals, powers of two (ties and absorbed terms), mixed magnitudes,
    zeros, and a negative / a huge term (those blocks take real additions)"""
    rng = np.random.default_rng(5)

    def plain(x, init):
        s = np.float64(init)
        for v in x:
            s = s + v
        return float(s)

    for trial in range(6):
        n = int(rng.integers(9000, 60000))
        if trial == 0:
            x = rng.random(n)
        elif trial == 1:
            s_ = rng.integers(1, 21, n)
            x = s_ / (40 - s_)
        elif trial == 2:
            x = np.ldexp(rng.integers(1, 8, n).astype(float), rng.integers(-60, 3, n))
        elif trial == 3:
            x = rng.random(n) * 10.0 ** rng.integers(-12, 6, n)
        elif trial == 4:
            x = np.ones(n)
            x[rng.integers(0, n, 50)] = 0.0
        else:
            x = rng.random(n)
            x[n // 2] = -0.3
            x[n // 3] = 1e18
        init = 0.0 if trial % 2 else float(rng.random() * 1000)
        assert emul_modopt.seq_sum(x, init) == plain(x, init), trial


def _write_edges(path, n1, n2, w=None):
    with open(path, "w") as f:
        for e in range(len(n1)):
            f.write(f"{n1[e]}\t{n2[e]}" + ("" if w is None else f"\t{float(w[e])!r}") + "\n")


@needs_ref
@pytest.mark.parametrize("seed", range(5))
def test_emulated_edge_list_input_matches_the_reference_reading_the_same_file(seed, tmp_path):
    """edge.file.name (RModularityOptimizer.cpp:55-63): the reference reads the file with its own readInputFile;
    the device steps get the parsed list (build_network_edges).  Lists in arbitrary order, with reversed pairs,
    self pairs and repeated pairs (only node1 < node2 counts, repeats stay separate edges)"""
    rng = np.random.default_rng(77 + seed)
    n = int(rng.integers(30, 600))
    m = int(n * rng.uniform(2, 9))
    n1 = rng.integers(0, n, m)
    n2 = rng.integers(0, n, m)
    if seed % 2 == 0:                       # the format WriteEdgeFile produces: col < row, column-major order
        lo, hi = np.minimum(n1, n2), np.maximum(n1, n2)
        keep = lo < hi
        order = np.lexsort((hi[keep], lo[keep]))
        n1, n2 = lo[keep][order], hi[keep][order]
    w = None if seed == 1 else np.round(rng.random(len(n1)) + 0.01, 6)
    path = str(tmp_path / "edges.txt")
    _write_edges(path, n1, n2, w)
    n_nodes = int(max(n1.max(), n2.max())) + 1
    for alg in (1, 2, 3):
        params = (1 + seed % 2, 0.8 if seed % 2 == 0 else 0.5, alg, 2, 4, 12345 + seed)
        want, wc, wq = oracle.modularity_cluster_file_ref(path, *params)
        assert want.size == n_nodes
        got, gc, gq = emul_modopt.cluster_edges(n1, n2, np.ones(len(n1)) if w is None else w, n_nodes, *params)
        assert gc == wc and (got == want).all() and gq == wq, (alg, params)


@needs_ref
def test_edge_file_round_trip_of_an_snn_graph(tmp_path):
    """WriteEdgeFile -> edge.file.name gives what the matrix input gives when the weights survive the
    15-digit text form (Jaccard values of k = 10: s / (20 - s))"""
    X = np.random.default_rng(3).normal(size=(700, 8))
    p, i, x = oracle.find_neighbors(X, k=10, prune=1 / 15)["snn"]
    path = str(tmp_path / "snn_edges.txt")
    oracle.write_edge_file(p, i, x, path)
    a, ac, aq = oracle.modularity_cluster_ref(p, i, x, 1, 0.8, 1, 3, 5, 9)
    b, bc, bq = oracle.modularity_cluster_file_ref(path, 1, 0.8, 1, 3, 5, 9)
    assert b.size <= a.size
    e1, e2, ew = np.loadtxt(path, delimiter="\t", unpack=True)
    got, gc, gq = emul_modopt.cluster_edges(e1.astype(np.int32), e2.astype(np.int32), ew, b.size, 1, 0.8, 1, 3, 5, 9)
    assert gc == bc and (got == b).all() and gq == bq


def test_r_random_draw_of_group_singletons_is_reproduced():
    """GroupSingletons breaks ties with set.seed(1); sample(tied, 1) (R/clustering.R:1395-1405).  The restated
    generator gives R's well-known outputs: set.seed(1); runif(6) and sample(1:10) under seeds 1 and 42"""
    from seurat_b200 import clustering as cl
    g = cl._r_unif_rand(1)
    assert [round(next(g), 7) for _ in range(6)] == [0.2655087, 0.3721239, 0.5728534, 0.9082078, 0.2016819, 0.8983897]
    assert [v + 1 for v in cl._r_sample(10, seed=1)] == [9, 4, 7, 1, 2, 5, 3, 10, 6, 8]
    assert [v + 1 for v in cl._r_sample(10, seed=42)] == [1, 5, 10, 8, 2, 4, 6, 9, 7, 3]
    assert [cl._r_sample(L, 1)[0] for L in range(1, 10)] == [0, 0, 0, 0, 0, 0, 0, 0, 8]
    # a singleton equally connected to two clusters joins the first of them (R's draw for two ties)
    import seurat_b200 as sb
    import scipy.sparse as sp
    S = sp.csc_matrix(np.array([[1, .5, 0, 0, .2], [.5, 1, 0, 0, .2], [0, 0, 1, .5, .2], [0, 0, .5, 1, .2],
                                [.2, .2, .2, .2, 1]]))
    assert sb.GroupSingletons([0, 0, 1, 1, 2], S, verbose=False) == [0, 0, 1, 1, 0]
