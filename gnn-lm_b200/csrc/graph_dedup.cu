// `--deprecated` graph assembly on the device: one ntgt node per DISTINCT datastore row of a block.
//
// Replaces GraphTokenBlockDataset.deprecated_build_graph (reference: fairseq/data/token_block_dataset.py:414-479):
// walking tokens, neighbours and their context rows in order, a row gets a node id the first time it is seen
// (`offsets2ntgt_id`, :440-448,461-466); ntgt-ntgt edges come from ONE build_ntgt_edges(context=1, bidirect=True) call over
// all nodes of the block (:469), i.e. rows at distance 1 are connected wherever they came from, every node has a self
// loop, and in insertion order the in-edges of the node of row o are [node(o-1), itself, node(o+1)] (forward edges
// first, reversed ones appended by the bidirect pass -- the order a stable sort by destination keeps).  Blocks of a batch
// do not share nodes (dgl.batch, fairseq/data/monolingual_dataset.py:261).
//
// The Python dict becomes an open-addressing hash table keyed by (block, row) that keeps the MINIMUM candidate position
// per key (atomicMin): candidate positions p = ((token * k + neighbour) * w + slot) enumerate the reference's visiting
// order, so "first time seen" == "the candidate whose position is the key's minimum", and node ids are the exclusive scan
// of those head flags.  No sort, no host synchronisation; the node count stays on the device.
//   insert -> flag heads -> scan -> assign ids -> neighbours / in-degree -> scan -> fill CSR, inter edges
#include "common.cuh"
#include "graph_shape.cuh"

namespace gnnlm {

constexpr int DD_THREADS = 256;
constexpr int DD_ITEMS = 4;
constexpr int DD_TILE = DD_THREADS * DD_ITEMS;
constexpr int64_t DD_EMPTY = -1;

__device__ __forceinline__ uint64_t dd_hash(int64_t key) {
  uint64_t x = (uint64_t)key;
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27; x *= 0x94d049bb133111ebULL;
  x ^= x >> 31;
  return x;
}

// candidate p -> (valid, key = block * n_datastore + row)
__device__ __forceinline__ bool dd_candidate(int64_t p, const int64_t* __restrict__ nbr, const int64_t* __restrict__ tgt_pos,
                                             int64_t L, int64_t k, int w, int64_t n_datastore, int left_ctx, int right_ctx,
                                             int64_t invalid_ctx, int64_t* key) {
  const int64_t c = p / w;
  const int s = (int)(p - c * w);
  const int64_t t = c / k;
  const int64_t o = __ldg(nbr + c);
  const ClusterShape sh = cluster_shape(o, tgt_pos ? __ldg(tgt_pos + t) : 0, n_datastore, left_ctx, right_ctx, invalid_ctx);
  if (!sh.valid || s > sh.nl + sh.nr) return false;
  // creation order inside a cluster: centre, left context ascending, right context ascending (:440,452-466)
  const int64_t row = s == 0 ? o : (s <= sh.nl ? o - sh.nl + (s - 1) : o + (s - sh.nl));
  *key = (t / L) * n_datastore + row;
  return true;
}

__global__ void __launch_bounds__(DD_THREADS) dd_insert_kernel(const int64_t* __restrict__ nbr, const int64_t* __restrict__ tgt_pos,
                                                               int64_t cap, int64_t L, int64_t k, int w, int64_t n_datastore,
                                                               int left_ctx, int right_ctx, int64_t invalid_ctx,
                                                               int64_t* __restrict__ keys, uint32_t* __restrict__ minpos,
                                                               uint64_t mask, int32_t* __restrict__ cand_slot) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= cap) return;
  int64_t key;
  if (!dd_candidate(p, nbr, tgt_pos, L, k, w, n_datastore, left_ctx, right_ctx, invalid_ctx, &key)) {
    cand_slot[p] = -1;
    return;
  }
  uint64_t h = dd_hash(key) & mask;
  while (true) {
    const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(keys + h), (unsigned long long)DD_EMPTY,
                                              (unsigned long long)key);
    if (prev == (unsigned long long)DD_EMPTY || prev == (unsigned long long)key) break;
    h = (h + 1) & mask;
  }
  atomicMin(minpos + h, (uint32_t)p);
  cand_slot[p] = (int32_t)h;
}

__global__ void __launch_bounds__(DD_THREADS) dd_flag_kernel(int64_t cap, const int32_t* __restrict__ cand_slot,
                                                             const uint32_t* __restrict__ minpos, int32_t* __restrict__ flag) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= cap) return;
  const int32_t h = cand_slot[p];
  flag[p] = (h >= 0 && minpos[h] == (uint32_t)p) ? 1 : 0;
}

// ---- exclusive scan of an int32 array: out[i] = sum_{j<i} in[j], out[n] = total (three passes, tile sums in ws)
__device__ __forceinline__ int dd_block_scan(int v, int* total) {
  __shared__ int warp_tot[DD_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int x = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += x;
  }
  if (lane == 31) warp_tot[wid] = v;
  __syncthreads();
  if (wid == 0) {
    int t = lane < DD_THREADS / 32 ? warp_tot[lane] : 0;
#pragma unroll
    for (int o = 1; o < DD_THREADS / 32; o <<= 1) {
      const int x = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += x;
    }
    if (lane < DD_THREADS / 32) warp_tot[lane] = t;
  }
  __syncthreads();
  if (wid > 0) v += warp_tot[wid - 1];
  *total = warp_tot[DD_THREADS / 32 - 1];
  __syncthreads();
  return v;
}

__global__ void __launch_bounds__(DD_THREADS) dd_scan_tiles(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ tile_sums) {
  const int64_t base = (int64_t)blockIdx.x * DD_TILE + threadIdx.x * DD_ITEMS;
  int acc = 0;
#pragma unroll
  for (int i = 0; i < DD_ITEMS; ++i)
    if (base + i < n) acc += in[base + i];
  int total;
  dd_block_scan(acc, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(DD_THREADS) dd_scan_sums(int32_t* __restrict__ tile_sums, int64_t n_tiles) {
  int carry = 0;
  for (int64_t c = 0; c < n_tiles; c += DD_THREADS) {
    const int64_t i = c + threadIdx.x;
    const int v = i < n_tiles ? tile_sums[i] : 0;
    int total;
    const int inc = dd_block_scan(v, &total);
    if (i < n_tiles) tile_sums[i] = carry + inc - v;
    carry += total;
  }
}

__global__ void __launch_bounds__(DD_THREADS) dd_scan_write(const int32_t* __restrict__ in, int64_t n,
                                                            const int32_t* __restrict__ tile_sums, int32_t* __restrict__ out) {
  const int64_t base = (int64_t)blockIdx.x * DD_TILE + threadIdx.x * DD_ITEMS;
  int item[DD_ITEMS], acc = 0;
#pragma unroll
  for (int i = 0; i < DD_ITEMS; ++i) {
    item[i] = base + i < n ? in[base + i] : 0;
    acc += item[i];
  }
  int total;
  const int inc = dd_block_scan(acc, &total);
  int run = tile_sums[blockIdx.x] + inc - acc;
#pragma unroll
  for (int i = 0; i < DD_ITEMS; ++i) {
    const int64_t idx = base + i;
    if (idx < n) out[idx] = run;
    run += item[i];
    if (idx == n - 1) out[n] = run;
  }
}

__global__ void __launch_bounds__(DD_THREADS) dd_assign_kernel(const int64_t* __restrict__ nbr, const int64_t* __restrict__ tgt_pos,
                                                               int64_t cap, int64_t L, int64_t k, int w, int64_t n_datastore,
                                                               int left_ctx, int right_ctx, int64_t invalid_ctx,
                                                               const int32_t* __restrict__ cand_slot,
                                                               const int32_t* __restrict__ flag, const int32_t* __restrict__ ids,
                                                               int32_t* __restrict__ slot_id, int64_t* __restrict__ node_key,
                                                               int64_t* __restrict__ ntgt_row, int32_t* __restrict__ n_ntgt) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p == 0) n_ntgt[0] = ids[cap];
  if (p >= cap || !flag[p]) return;
  int64_t key;
  dd_candidate(p, nbr, tgt_pos, L, k, w, n_datastore, left_ctx, right_ctx, invalid_ctx, &key);
  const int32_t id = ids[p];
  slot_id[cand_slot[p]] = id;
  node_key[id] = key;
  ntgt_row[id] = key % n_datastore;
}

__device__ __forceinline__ int32_t dd_lookup(int64_t key, const int64_t* __restrict__ keys, const int32_t* __restrict__ slot_id,
                                             uint64_t mask) {
  uint64_t h = dd_hash(key) & mask;
  while (true) {
    const int64_t kk = keys[h];
    if (kk == key) return slot_id[h];
    if (kk == DD_EMPTY) return -1;
    h = (h + 1) & mask;
  }
}

// in-degree and the ids of the nodes of rows o-1 / o+1 of the same block; entries beyond the live count get degree 0
__global__ void __launch_bounds__(DD_THREADS) dd_degree_kernel(int64_t cap, const int32_t* __restrict__ n_ntgt,
                                                               const int64_t* __restrict__ node_key, int64_t n_datastore,
                                                               const int64_t* __restrict__ keys,
                                                               const int32_t* __restrict__ slot_id, uint64_t mask,
                                                               int32_t* __restrict__ left_id, int32_t* __restrict__ right_id,
                                                               int32_t* __restrict__ deg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  if (i >= n_ntgt[0]) {
    deg[i] = 0;
    return;
  }
  const int64_t key = node_key[i];
  const int64_t row = key % n_datastore;
  const int32_t l = row > 0 ? dd_lookup(key - 1, keys, slot_id, mask) : -1;
  const int32_t r = row + 1 < n_datastore ? dd_lookup(key + 1, keys, slot_id, mask) : -1;
  left_id[i] = l;
  right_id[i] = r;
  deg[i] = 1 + (l >= 0) + (r >= 0);
}

__global__ void __launch_bounds__(DD_THREADS) dd_fill_kernel(int64_t cap, const int32_t* __restrict__ n_ntgt,
                                                             const int32_t* __restrict__ left_id,
                                                             const int32_t* __restrict__ right_id,
                                                             const int32_t* __restrict__ indptr, int32_t* __restrict__ indices) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap || i >= n_ntgt[0]) return;
  int32_t e = indptr[i];
  if (left_id[i] >= 0) indices[e++] = left_id[i];
  indices[e++] = (int32_t)i;
  if (right_id[i] >= 0) indices[e++] = right_id[i];
}

// ('ntgt','inter','tgt'): the centre node of every valid (token, neighbour) pair, in (token, neighbour) order (:449-450)
__global__ void __launch_bounds__(DD_THREADS) dd_inter_kernel(int64_t T, int64_t k, int w, const int32_t* __restrict__ valid_base,
                                                              const int32_t* __restrict__ cand_slot,
                                                              const int32_t* __restrict__ slot_id,
                                                              int32_t* __restrict__ inter_indptr,
                                                              int32_t* __restrict__ inter_indices) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c <= T) inter_indptr[c] = valid_base[c * k];
  if (c >= T * k) return;
  const int32_t h = cand_slot[c * w];
  if (h >= 0) inter_indices[valid_base[c]] = slot_id[h];
}

struct DedupLayout {
  int64_t hcap, cap, n_tiles;
  size_t off_keys, off_minpos, off_slot_id, off_cand, off_flag, off_ids, off_tiles, off_node_key, off_left, off_right, off_deg,
      total;
};

static DedupLayout dedup_layout(int64_t T, int64_t k, int32_t w) {
  DedupLayout l;
  l.cap = T * k * w;
  l.hcap = 1024;
  while (l.hcap < 2 * l.cap) l.hcap <<= 1;
  l.n_tiles = ceil_div(l.cap + 1, DD_TILE);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
  l.off_keys = take((size_t)l.hcap * 8);
  l.off_minpos = take((size_t)l.hcap * 4);          // directly after the keys: one memset(0xFF) initialises both
  l.off_slot_id = take((size_t)l.hcap * 4);
  l.off_cand = take((size_t)l.cap * 4);
  l.off_flag = take((size_t)(l.cap + 1) * 4);
  l.off_ids = take((size_t)(l.cap + 1) * 4);
  l.off_tiles = take((size_t)l.n_tiles * 4);
  l.off_node_key = take((size_t)l.cap * 8);
  l.off_left = take((size_t)l.cap * 4);
  l.off_right = take((size_t)l.cap * 4);
  l.off_deg = take((size_t)(l.cap + 1) * 4);
  l.total = o;
  return l;
}

}  // namespace gnnlm

using namespace gnnlm;

extern "C" int64_t gnnlm_graph_dedup_workspace_bytes(int64_t T, int64_t k, int32_t w) {
  if (T <= 0 || k <= 0 || w <= 0) return 0;
  return (int64_t)dedup_layout(T, k, w).total;
}

extern "C" int32_t gnnlm_graph_dedup(const int64_t* nbr, const int64_t* tgt_pos, int64_t T, int64_t L, int64_t k,
                                     int64_t n_datastore, int32_t left_ctx, int32_t right_ctx, int64_t invalid_ctx,
                                     const int32_t* valid_base, int64_t* ntgt_row, int32_t* n_ntgt, int32_t* nn_indptr,
                                     int32_t* nn_indices, int32_t* inter_indptr, int32_t* inter_indices, void* workspace,
                                     int64_t workspace_bytes, gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(nbr && valid_base && ntgt_row && n_ntgt && nn_indptr && nn_indices && inter_indptr && inter_indices && workspace,
                  GNNLM_E_ARG, "gnnlm_graph_dedup: null pointer");
  GNNLM_CHECK_ARG(T > 0 && L > 0 && T % L == 0 && k > 0 && n_datastore > 0 && left_ctx >= 0 && right_ctx >= 0, GNNLM_E_SHAPE,
                  "gnnlm_graph_dedup: bad sizes (T must be a multiple of the block length L)");
  const int w = 1 + left_ctx + right_ctx;
  const DedupLayout l = dedup_layout(T, k, w);
  GNNLM_CHECK_ARG(l.cap < ((int64_t)1 << 31) - 1, GNNLM_E_SHAPE, "gnnlm_graph_dedup: too many candidate nodes for 32-bit ids");
  GNNLM_CHECK_ARG((T / L) < ((int64_t)1 << 62) / n_datastore, GNNLM_E_SHAPE, "gnnlm_graph_dedup: key overflow");
  GNNLM_CHECK_ARG(workspace_bytes >= (int64_t)l.total && (uintptr_t)workspace % 256 == 0, GNNLM_E_ARG,
                  "gnnlm_graph_dedup: workspace too small / not 256 B aligned (gnnlm_graph_dedup_workspace_bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = reinterpret_cast<char*>(workspace);
  int64_t* keys = reinterpret_cast<int64_t*>(ws + l.off_keys);
  uint32_t* minpos = reinterpret_cast<uint32_t*>(ws + l.off_minpos);
  int32_t* slot_id = reinterpret_cast<int32_t*>(ws + l.off_slot_id);
  int32_t* cand = reinterpret_cast<int32_t*>(ws + l.off_cand);
  int32_t* flag = reinterpret_cast<int32_t*>(ws + l.off_flag);
  int32_t* ids = reinterpret_cast<int32_t*>(ws + l.off_ids);
  int32_t* tiles = reinterpret_cast<int32_t*>(ws + l.off_tiles);
  int64_t* node_key = reinterpret_cast<int64_t*>(ws + l.off_node_key);
  int32_t* left = reinterpret_cast<int32_t*>(ws + l.off_left);
  int32_t* right = reinterpret_cast<int32_t*>(ws + l.off_right);
  int32_t* deg = reinterpret_cast<int32_t*>(ws + l.off_deg);
  const uint64_t mask = (uint64_t)l.hcap - 1;
  const unsigned g_cap = (unsigned)ceil_div(l.cap, DD_THREADS);
  const unsigned g_tiles = (unsigned)ceil_div(l.cap, DD_TILE);

  GNNLM_CUDA(cudaMemsetAsync(ws + l.off_keys, 0xFF, l.off_slot_id - l.off_keys, st));      // keys = -1, minpos = UINT_MAX
  dd_insert_kernel<<<g_cap, DD_THREADS, 0, st>>>(nbr, tgt_pos, l.cap, L, k, w, n_datastore, left_ctx, right_ctx, invalid_ctx,
                                                keys, minpos, mask, cand);
  dd_flag_kernel<<<g_cap, DD_THREADS, 0, st>>>(l.cap, cand, minpos, flag);
  dd_scan_tiles<<<g_tiles, DD_THREADS, 0, st>>>(flag, l.cap, tiles);
  dd_scan_sums<<<1, DD_THREADS, 0, st>>>(tiles, g_tiles);
  dd_scan_write<<<g_tiles, DD_THREADS, 0, st>>>(flag, l.cap, tiles, ids);
  dd_assign_kernel<<<g_cap, DD_THREADS, 0, st>>>(nbr, tgt_pos, l.cap, L, k, w, n_datastore, left_ctx, right_ctx, invalid_ctx,
                                                cand, flag, ids, slot_id, node_key, ntgt_row, n_ntgt);
  dd_degree_kernel<<<g_cap, DD_THREADS, 0, st>>>(l.cap, n_ntgt, node_key, n_datastore, keys, slot_id, mask, left, right, deg);
  dd_scan_tiles<<<g_tiles, DD_THREADS, 0, st>>>(deg, l.cap, tiles);
  dd_scan_sums<<<1, DD_THREADS, 0, st>>>(tiles, g_tiles);
  dd_scan_write<<<g_tiles, DD_THREADS, 0, st>>>(deg, l.cap, tiles, nn_indptr);
  dd_fill_kernel<<<g_cap, DD_THREADS, 0, st>>>(l.cap, n_ntgt, left, right, nn_indptr, nn_indices);
  dd_inter_kernel<<<(unsigned)ceil_div(T * k + 1, DD_THREADS), DD_THREADS, 0, st>>>(T, k, w, valid_base, cand, slot_id,
                                                                                  inter_indptr, inter_indices);
  GNNLM_LAUNCH_CHECK("gnnlm_graph_dedup");
  return 0;
}
