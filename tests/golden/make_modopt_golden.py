"""Golden vectors of the modularity optimiser, produced by the REAL reference
(/root/reference/src/ModularityOptimizer.cpp compiled into oracle/_ref/libmodopt_ref.so) in the
build container; /root/reference does not exist on the GPU box, so the vectors are committed.

    python tests/golden/make_modopt_golden.py            # -> tests/golden/modopt_fixtures.json

Before anything is written the reference build is checked against the known answers the
reference's own tests hold (tests/testthat/test_modularity_optimizer.R)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import modopt_cases as mc  # noqa: E402
import oracle  # noqa: E402


def snn_case(n, d, seed, centres, k):
    from seurat_b200.synthetic import gaussian_mixture
    X, _ = gaussian_mixture(n, d, seed=seed, n_centres=centres)
    r = oracle.find_neighbors(X, k=k, prune=1 / 15, method="kdtree")
    p, i, x = r["snn"]
    return p.astype(np.int32), i.astype(np.int32), x


def main():
    for (params, want) in mc.KARATE_KAT + mc.KARATE_KAT_SLM:
        got, _, _ = oracle.modularity_cluster_ref(*mc.slots(mc.karate()), *params)
        assert got.tolist() == want, f"oracle/_ref disagrees with the reference's own test vector for {params}"
    out = {"made_by": "tests/golden/make_modopt_golden.py with oracle/_ref/libmodopt_ref.so (the reference's own "
                      "ModularityOptimizer.cpp)", "cases": []}
    for name, build, plist in mc.small_cases():
        p, i, x = mc.slots(build())
        for params in plist:
            lab, nc, q = oracle.modularity_cluster_ref(p, i, x, *params)
            out["cases"].append({"graph": name, "params": list(params), "n_clusters": int(nc), "modularity": float(q).hex(),
                                 "sha256": mc.digest(lab), "labels": lab.tolist() if lab.size <= 1000 else lab[:64].tolist()})
            print(name, params, nc, q)
    for (n, d, seed, centres, k), plist in (((20000, 50, 7, 12, 20), [(1, 0.8, 1, 3, 10, 0), (1, 0.8, 2, 1, 10, 42),
                                                                       (2, 0.0001, 1, 1, 3, 1), (1, 0.8, 3, 1, 2, 0)]),
                                            ((100000, 50, 0, 32, 20), [(1, 0.8, 1, 2, 10, 0), (1, 0.8, 3, 1, 2, 0)])):
        p, i, x = snn_case(n, d, seed, centres, k)
        for params in plist:
            lab, nc, q = oracle.modularity_cluster_ref(p, i, x, *params)
            out["cases"].append({"graph": f"snn:{n}:{d}:{seed}:{centres}:{k}", "params": list(params), "n_clusters": int(nc),
                                 "modularity": float(q).hex(), "sha256": mc.digest(lab), "labels": lab[:64].tolist()})
            print("snn", n, params, nc, q)
    with open(os.path.join(HERE, "modopt_fixtures.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()
