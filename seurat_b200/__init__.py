"""seurat_b200 -- B200-native (sm_100a) implementation of Seurat's FindNeighbors
hot path: exact kNN (the nn.method="rann" arm of NNHelper) + nn Graph + SNN
(ComputeSNN), behind the reference's own interface, plus the modularity optimiser that
consumes the SNN graph (FindClusters, algorithms 1-3) and the SNN side consumers of src/snn.cpp.  See DESIGN.md."""
from ._lib import B200Error, Context, MultiContext, default_context, device_count  # noqa: F401
from .neighbors import (Cells, ComputeSNN, ComputeSNNwidth, DirectSNNToFile, Distances, Embedding, FindNeighbors,  # noqa: F401
                        FindNeighborsResident, FindNeighborsSeurat, Graph, Indices, L2Norm, Neighbor, NNdist, NNHelper, SeuratLite,
                        SNN_SmallestNonzero_Dist, WriteEdgeFile, fast_dist)

from .clustering import FindClusters, GroupSingletons, RunModularityClustering  # noqa: F401

__version__ = "0.1.0"
