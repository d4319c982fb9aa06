// knn_tc.cu -- tcgen05 candidate generation + FP64 re-rank (placeholder until
// the tensor path lands; the dispatcher falls back to the FP64 exact scan).
#include "internal.cuh"

bool knn_tc_applicable(int64_t, int64_t, int, int) { return false; }

int knn_tc_launch(b200nn_ctx *, const double *, int64_t, const double *, int64_t, int, int, int,
                  int32_t *, double *, b200nn_stats *) {
  b200nn_set_error("tensor-core kNN path not built");
  return B200NN_ERR_INVALID;
}
