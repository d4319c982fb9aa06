# b200nn.R -- R side of the drop-in (not run in the build container: no R there).
#
# Adds one arm to NNHelper's switch (R/clustering.R:1660-1677) and routes ComputeSNN
# (R/RcppExports.R:96-98) through the GPU library when options(Seurat.nn.backend = "b200").
# FindNeighbors.default / .Seurat keep their signatures and return values
# (list(nn = Graph, snn = Graph) or a Neighbor), see R/clustering.R:582-685, 790-898.

# cache.index / index (R/clustering.R:596-597, 1686-1688): the "index" of the exact search is
# the reference embedding kept on the GPU (an external pointer with a finaliser); NNHelper's
# unchanged tail stores results$idx in the Neighbor's alg.idx slot.
b200_nn2 <- function(data, query = data, k, cache.index = FALSE, index = NULL, ...) {
  # NNHelper passes query = data as the SAME object for the self search: identical() returns at
  # its pointer check then, without comparing 400 MB; only equal-shaped distinct matrices are compared
  self <- missing(query) || (identical(dim(query), dim(data)) && identical(query, data))
  if (!self) storage.mode(query) <- "double"
  if (is.null(index) && isTRUE(cache.index)) {
    storage.mode(data) <- "double"
    index <- .Call("b200_index", data, PACKAGE = "b200nn")
  }
  if (!is.null(index)) {
    res <- .Call("b200_index_nn2", index, if (self) NULL else query, as.integer(k), PACKAGE = "b200nn")
    if (isTRUE(cache.index)) res$idx <- index
    return(res)
  }
  storage.mode(data) <- "double"
  .Call("b200_nn2", data, if (self) NULL else query, as.integer(k), PACKAGE = "b200nn")
}

# In NNHelper (R/clustering.R:1660):
#   "b200" = {
#     args <- args[intersect(x = names(x = args), y = names(x = formals(fun = b200_nn2)))]
#     do.call(what = 'b200_nn2', args = args)
#   },
# and nn.method = "rann" is redirected there when the backend option is set:
#   if (method == "rann" && identical(getOption("Seurat.nn.backend"), "b200")) method <- "b200"

ComputeSNN <- function(nn_ranked, prune) {
  if (identical(getOption("Seurat.nn.backend"), "b200")) {
    return(.Call("b200_ComputeSNN", nn_ranked, as.double(prune), PACKAGE = "b200nn"))
  }
  .Call('_Seurat_ComputeSNN', PACKAGE = 'Seurat', nn_ranked, prune)
}

# Optional fused path inside FindNeighbors.default, replacing R/clustering.R:633-682 when
# query is NULL, return.neighbor is FALSE, nn.method is "rann"/"b200" and nn.eps == 0:
#   g <- .Call("b200_FindNeighbors", object, as.integer(k.param), as.double(prune.SNN),
#              isTRUE(compute.SNN), PACKAGE = "b200nn")
#   for (nm in names(g)) { dimnames(g[[nm]]) <- list(rownames(object), rownames(object))
#                          g[[nm]] <- as.Graph(g[[nm]]) }
#   return(g)

# The step after FindNeighbors and the SNN side consumers (R/RcppExports.R): same arguments, same values.
b200_backend <- function() identical(getOption("Seurat.nn.backend"), "b200")
RunModularityClusteringCpp <- function(SNN, modularityFunction, resolution, algorithm, nRandomStarts, nIterations,
                                       randomSeed, printOutput, edgefilename) {
  if (b200_backend())
    return(.Call("b200_RunModularityClusteringCpp", SNN, modularityFunction, resolution, algorithm, nRandomStarts,
                 nIterations, randomSeed, printOutput, edgefilename, PACKAGE = "b200nn"))
  .Call('_Seurat_RunModularityClusteringCpp', PACKAGE = 'Seurat', SNN, modularityFunction, resolution, algorithm,
        nRandomStarts, nIterations, randomSeed, printOutput, edgefilename)
}
WriteEdgeFile <- function(snn, filename, display_progress) {
  if (b200_backend()) return(invisible(.Call("b200_WriteEdgeFile", snn, filename, display_progress, PACKAGE = "b200nn")))
  invisible(.Call('_Seurat_WriteEdgeFile', PACKAGE = 'Seurat', snn, filename, display_progress))
}
DirectSNNToFile <- function(nn_ranked, prune, display_progress, filename) {
  if (b200_backend())
    return(.Call("b200_DirectSNNToFile", nn_ranked, as.double(prune), display_progress, filename, PACKAGE = "b200nn"))
  .Call('_Seurat_DirectSNNToFile', PACKAGE = 'Seurat', nn_ranked, prune, display_progress, filename)
}
fast_dist <- function(x, y, n) {
  if (b200_backend()) return(.Call("b200_fast_dist", x, y, n, PACKAGE = "b200nn"))
  .Call('_Seurat_fast_dist', PACKAGE = 'Seurat', x, y, n)
}
SNN_SmallestNonzero_Dist <- function(snn, mat, n, nearest_dist) {
  if (b200_backend())
    return(.Call("b200_SNN_SmallestNonzero_Dist", snn, mat, as.integer(n), nearest_dist, PACKAGE = "b200nn"))
  .Call('_Seurat_SNN_SmallestNonzero_Dist', PACKAGE = 'Seurat', snn, mat, n, nearest_dist)
}

.onUnload <- function(libpath) .Call("b200_release", PACKAGE = "b200nn")
