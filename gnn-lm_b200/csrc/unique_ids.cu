// Distinct centre rows of a batch: the ntgt side of a cluster is a pure function of its centre's datastore row.
//
// new_build_graph never de-duplicates (fairseq/data/token_block_dataset.py:355 "todo", :363-374): two (token, neighbour) pairs
// that retrieved the same datastore row o get two clusters with the same rows o - c_l .. o + c_r, the same chain edges and
// therefore -- ntgt nodes only ever receive messages from inside their cluster (hgt.py:350-358 over ('ntgt','intra','ntgt')) --
// the same features in every layer.  Real kNN graphs repeat centres (the same context recurs inside a block); computing each
// distinct centre once and gathering its features back gives results identical to the reference's duplicated clusters.
//
// gnnlm_unique_centres: open-addressing hash table over the valid pairs' ids (first inserter claims a compact index by
// atomicAdd), then a lookup pass -> uniq [n] (distinct ids, -1 padded: a neighbour array for gnnlm_graph_count / _fill with
// k = 1) and inv [n_valid] (compact valid pair -> index into uniq).  The ORDER of uniq depends on the race and differs from
// run to run; per-row results do not (every kernel downstream computes a row from that row's inputs only).
#include "common.cuh"

namespace gnnlm {

__device__ __forceinline__ uint32_t uq_hash(int64_t key, int log2cap) {
  return (uint32_t)(((uint64_t)key * 0x9E3779B97F4A7C15ull) >> (64 - log2cap));
}

__global__ void __launch_bounds__(256) unique_insert_kernel(const int64_t* __restrict__ nbr, const int32_t* __restrict__ valid_base,
                                                            int64_t n, long long* __restrict__ keys, int32_t* __restrict__ slot_idx,
                                                            int log2cap, int64_t* __restrict__ uniq, int32_t* __restrict__ counter) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || valid_base[i + 1] == valid_base[i]) return;
  const long long key = nbr[i];
  const uint32_t mask = (1u << log2cap) - 1;
  for (uint32_t s = uq_hash(key, log2cap);; s = (s + 1) & mask) {
    const long long prev = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(keys + s), (unsigned long long)-1ll,
                                                (unsigned long long)key);
    if (prev == -1ll) {                                     // claimed: this id is new
      const int32_t idx = atomicAdd(counter, 1);
      uniq[idx] = key;
      slot_idx[s] = idx;
      return;
    }
    if (prev == key) return;
  }
}

__global__ void __launch_bounds__(256) unique_lookup_kernel(const int64_t* __restrict__ nbr, const int32_t* __restrict__ valid_base,
                                                            int64_t n, const long long* __restrict__ keys,
                                                            const int32_t* __restrict__ slot_idx, int log2cap,
                                                            int32_t* __restrict__ inv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t vb = valid_base[i];
  if (valid_base[i + 1] == vb) return;
  const long long key = nbr[i];
  const uint32_t mask = (1u << log2cap) - 1;
  uint32_t s = uq_hash(key, log2cap);
  while (keys[s] != key) s = (s + 1) & mask;                // present by construction
  inv[vb] = slot_idx[s];
}

static int uq_log2cap(int64_t n) {
  int l = 4;
  while ((1ll << l) < 2 * n) ++l;
  return l;
}

}  // namespace gnnlm

using namespace gnnlm;

extern "C" int64_t gnnlm_unique_workspace_bytes(int64_t n) {
  if (n <= 0) return 16;
  return (int64_t)(sizeof(long long) + sizeof(int32_t)) << uq_log2cap(n);
}

extern "C" int32_t gnnlm_unique_centres(const int64_t* nbr, const int32_t* valid_base, int64_t n, int64_t* uniq, int32_t* inv,
                                        int32_t* n_unique, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  GNNLM_CHECK_ARG(nbr && valid_base && uniq && inv && n_unique && workspace, GNNLM_E_ARG, "gnnlm_unique_centres: null pointer");
  GNNLM_CHECK_ARG(n >= 0 && n < (1ll << 30), GNNLM_E_SHAPE, "gnnlm_unique_centres: bad size");
  GNNLM_CHECK_ARG(workspace_bytes >= gnnlm_unique_workspace_bytes(n), GNNLM_E_SHAPE, "gnnlm_unique_centres: workspace too small");
  GNNLM_CUDA(cudaMemsetAsync(n_unique, 0, sizeof(int32_t), stream));
  if (n == 0) return 0;
  const int l2 = uq_log2cap(n);
  long long* keys = reinterpret_cast<long long*>(workspace);
  int32_t* slot_idx = reinterpret_cast<int32_t*>(keys + (1ll << l2));
  GNNLM_CUDA(cudaMemsetAsync(keys, 0xFF, sizeof(long long) << l2, stream));          // -1 = empty (valid ids are >= 0)
  GNNLM_CUDA(cudaMemsetAsync(uniq, 0xFF, sizeof(int64_t) * n, stream));              // -1 = no neighbour
  const unsigned grid = (unsigned)ceil_div(n, 256);
  unique_insert_kernel<<<grid, 256, 0, stream>>>(nbr, valid_base, n, keys, slot_idx, l2, uniq, n_unique);
  GNNLM_LAUNCH_CHECK("gnnlm_unique_centres(insert)");
  unique_lookup_kernel<<<grid, 256, 0, stream>>>(nbr, valid_base, n, keys, slot_idx, l2, inv);
  GNNLM_LAUNCH_CHECK("gnnlm_unique_centres(lookup)");
  return 0;
}
