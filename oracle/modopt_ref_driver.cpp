// modopt_ref_driver.cpp -- C entry point over the REAL reference modularity optimiser.
//
// TEST INFRASTRUCTURE ONLY (like everything under oracle/).  This file holds no algorithm: it is
// compiled together with /root/reference/src/ModularityOptimizer.cpp (read where it lies, never
// copied) into oracle/_ref/libmodopt_ref.so by `make -C oracle ref`, and only drives the reference
// classes the way the Rcpp wrapper does:
//
//   /root/reference/src/RModularityOptimizer.cpp:21-173  RunModularityClusteringCpp
//     :72-81    lower triangle of the SNN dgCMatrix -> (node1 = col, node2 = row, weight)
//     :85-86    matrixToNetwork(node1, node2, weights, modularityFunction, nNodes)
//     :98       resolution2 (standard modularity: resolution / (2 m + self links))
//     :103-146  random starts x iterations of Louvain / Louvain + refinement / SLM
//     :165-167  orderClustersByNNodes, cluster vector out
//
// The wrapper itself needs Rcpp + RcppEigen (absent from the image), so its ~60 lines of glue are
// restated here against the same public classes; everything numerical runs in the reference's
// own object code.  kind = "reference" in bench.py's cpu_baseline for the clustering step.
#include <cstdint>
#include <limits>
#include <memory>

#include "ModularityOptimizer.h"

using namespace ModularityOptimizer;

// RModularityOptimizer.cpp:98-167 on a loaded network
static int run_on(std::shared_ptr<Network> network, int n, int modularityFunction, double resolution, int algorithm,
                  int nRandomStarts, int nIterations, uint64_t randomSeed, int *cluster_out, double *modularity_out) {
    const double resolution2 =
        modularityFunction == 1
            ? resolution / (2 * network->getTotalEdgeWeight() + network->getTotalEdgeWeightSelfLinks())
            : resolution;
    std::shared_ptr<Clustering> clustering;
    double maxModularity = -std::numeric_limits<double>::infinity(), modularity = 0;
    JavaRandom random(randomSeed);
    for (int s = 0; s < nRandomStarts; ++s) {
      VOSClusteringTechnique t(network, resolution2);
      int j = 0;
      bool update = true;
      do {
        if (algorithm == 1)
          update = t.runLouvainAlgorithm(random);
        else if (algorithm == 2)
          update = t.runLouvainAlgorithmWithMultilevelRefinement(random);
        else if (algorithm == 3)
          t.runSmartLocalMovingAlgorithm(random);
        j++;
        modularity = t.calcQualityFunction();
      } while (j < nIterations && update);
      if (modularity > maxModularity) {
        clustering = t.getClustering();
        maxModularity = modularity;
      }
    }
    if (!clustering) return -3;
    clustering->orderClustersByNNodes();
    for (int v = 0; v < n; ++v) cluster_out[v] = clustering->cluster[v];
    if (modularity_out) *modularity_out = maxModularity;
    return clustering->getNClusters();
}

extern "C" {

// p/i/x: dgCMatrix slots of the (symmetric) SNN, n columns.  cluster_out[n].  Returns the number
// of clusters, or a negative code: -1 invalid argument (the wrapper's stop() cases), -2 no edges
// ("Matrix contained no network data"), -3 clustering failed, -4 C++ exception.
int ref_modularity_cluster(int n, const int *p, const int *i, const double *x, int modularityFunction,
                           double resolution, int algorithm, int nRandomStarts, int nIterations,
                           uint64_t randomSeed, int *cluster_out, double *modularity_out) {
  if (modularityFunction != 1 && modularityFunction != 2) return -1;
  if (algorithm < 1 || algorithm > 4) return -1;
  if (nRandomStarts < 1 || nIterations < 1) return -1;
  if (modularityFunction == 2 && resolution > 1.0) return -1;
  try {
    IVector node1, node2;
    DVector w;
    for (int c = 0; c < n; ++c)
      for (int e = p[c]; e < p[c + 1]; ++e) {
        if (c >= i[e]) continue;  // it.col() >= it.row(): keep the strict lower triangle
        node1.emplace_back(c);
        node2.emplace_back(i[e]);
        w.emplace_back(x[e]);
      }
    if (node1.empty()) return -2;
    std::shared_ptr<Network> network = matrixToNetwork(node1, node2, w, modularityFunction, n);
    return run_on(network, n, modularityFunction, resolution, algorithm, nRandomStarts, nIterations, randomSeed,
                  cluster_out, modularity_out);
  } catch (...) {
    return -4;
  }
}

// The edge.file.name input (RModularityOptimizer.cpp:55-63): the reference's own readInputFile
// (ModularityOptimizer.cpp:806-838).  *n_nodes_out = number of nodes; cluster_out needs that many
// entries (capacity); -5 when the file could not be read ("Could not parse edge file.").
int ref_modularity_cluster_file(const char *fname, int modularityFunction, double resolution, int algorithm,
                                int nRandomStarts, int nIterations, uint64_t randomSeed, int *cluster_out,
                                int capacity, int *n_nodes_out, double *modularity_out) {
  if (modularityFunction != 1 && modularityFunction != 2) return -1;
  if (algorithm < 1 || algorithm > 4) return -1;
  if (nRandomStarts < 1 || nIterations < 1) return -1;
  if (modularityFunction == 2 && resolution > 1.0) return -1;
  std::shared_ptr<Network> network;
  try {
    network = readInputFile(fname, modularityFunction);
  } catch (...) {
    return -5;
  }
  try {
    const int n = network->getNNodes();
    if (n_nodes_out) *n_nodes_out = n;
    if (n > capacity) return -1;
    return run_on(network, n, modularityFunction, resolution, algorithm, nRandomStarts, nIterations, randomSeed,
                  cluster_out, modularity_out);
  } catch (...) {
    return -4;
  }
}

}  // extern "C"
