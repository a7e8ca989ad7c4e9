// Shape of the context cluster of one (token, neighbour) pair -- shared by the two graph builders.
#pragma once
#include <stdint.h>

namespace gnnlm {

struct ClusterShape {
  int nl, nr, valid;
};

__device__ __forceinline__ ClusterShape cluster_shape(int64_t o, int64_t pos, int64_t n_datastore, int left_ctx,
                                                      int right_ctx, int64_t invalid_ctx) {
  ClusterShape s{0, 0, 0};
  if (o == -1) return s;                                       // token_block_dataset.py:358
  // ids outside the datastore never reach the code / label gathers (the host side raises IndexError before a block with
  // such ids is staged, dataset.GraphTokenBlockDataset.__getitem__; callers that bypass it get an absent neighbour)
  if (o < 0 || o >= n_datastore) return s;
  if (invalid_ctx > 0) {
    int64_t dlt = pos - o;
    if (dlt < 0) dlt = -dlt;
    if (dlt < invalid_ctx) return s;                           // :361
  }
  s.valid = 1;
  // left: range(max(0, o - c_l), o)            (:380)
  int64_t lo = o - left_ctx;
  if (lo < 0) lo = 0;
  int64_t nl = o - lo;
  s.nl = nl > 0 ? (int)nl : 0;
  // right: range(o + 1, min(N, o + 1 + c_r))   (:384, with the Q1 fix)
  int64_t hi = o + 1 + right_ctx;
  if (hi > n_datastore) hi = n_datastore;
  int64_t nr = hi - (o + 1);
  s.nr = nr > 0 ? (int)nr : 0;
  return s;
}

}  // namespace gnnlm
