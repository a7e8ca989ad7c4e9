// Graph assembly on the device: neighbour ids -> ntgt node table + canonical CSRs.
//
// Replaces the pure-Python triple loop of GraphTokenBlockDataset.new_build_graph
// (reference: fairseq/data/token_block_dataset.py:338-412) together with build_ntgt_edges
// (:545-584), auto_regressive_edges (:586-594) and the node-id offsets of dgl.batch
// (fairseq/data/monolingual_dataset.py:261).
//
// Observation that makes this a count -> scan -> fill problem: new_build_graph never de-duplicates
// (:355 "todo"), so every valid (token, neighbour) pair creates one *cluster* of contiguous
// datastore rows [o - nl, o + nr], nl = min(c_l, o), nr = clip(min(N, o+1+c_r) - (o+1)), whose nodes
// are numbered centre, left-ascending, right-ascending, and build_ntgt_edges(context=1,
// bidirect=True) on contiguous rows is a chain with self loops.  In insertion order the in-edges
// of the node at sorted position p are [p-1, p, p+1] (forward edges first, reversed ones appended
// by the bidirect pass), which is exactly what a stable sort by destination preserves.
//
// HBM-bound integer work: 8 B read per (token, neighbour), ~(8+4+4) B + 4*(3w-2)/w B written per
// node.  One thread per cluster; scans are block-local + a single-block pass over block sums.
#include "common.cuh"
#include "graph_shape.cuh"

namespace gnnlm {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int2 add2(int2 a, int2 b) { return make_int2(a.x + b.x, a.y + b.y); }

// block-wide inclusive scan of one int2 per thread; returns inclusive value, total in *total
__device__ __forceinline__ int2 block_scan_inclusive(int2 v, int2* total) {
  __shared__ int2 warp_tot[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int x = __shfl_up_sync(0xffffffffu, v.x, o);
    int y = __shfl_up_sync(0xffffffffu, v.y, o);
    if (lane >= o) { v.x += x; v.y += y; }
  }
  if (lane == 31) warp_tot[wid] = v;
  __syncthreads();
  if (wid == 0) {
    int2 t = lane < SCAN_THREADS / 32 ? warp_tot[lane] : make_int2(0, 0);
#pragma unroll
    for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
      int x = __shfl_up_sync(0xffffffffu, t.x, o);
      int y = __shfl_up_sync(0xffffffffu, t.y, o);
      if (lane >= o) { t.x += x; t.y += y; }
    }
    if (lane < SCAN_THREADS / 32) warp_tot[lane] = t;
  }
  __syncthreads();
  if (wid > 0) v = add2(v, warp_tot[wid - 1]);
  *total = warp_tot[SCAN_THREADS / 32 - 1];
  __syncthreads();
  return v;
}

// pass A: per-tile totals of (nodes, valid)
__global__ void __launch_bounds__(SCAN_THREADS) graph_count_tiles(const int64_t* __restrict__ nbr,
                                                                  const int64_t* __restrict__ tgt_pos, int64_t n, int64_t k,
                                                                  int64_t n_datastore, int left_ctx, int right_ctx,
                                                                  int64_t invalid_ctx, int2* __restrict__ tile_sums) {
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int2 acc = make_int2(0, 0);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t idx = base + i;
    if (idx < n) {
      int64_t pos = tgt_pos ? __ldg(tgt_pos + idx / k) : 0;
      ClusterShape s = cluster_shape(__ldg(nbr + idx), pos, n_datastore, left_ctx, right_ctx, invalid_ctx);
      acc.x += s.valid ? 1 + s.nl + s.nr : 0;
      acc.y += s.valid;
    }
  }
  int2 total;
  block_scan_inclusive(acc, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// pass B: exclusive scan of tile sums, single block, sequential over chunks of SCAN_THREADS
__global__ void __launch_bounds__(SCAN_THREADS) graph_scan_tiles(int2* __restrict__ tile_sums, int64_t n_tiles) {
  int2 carry = make_int2(0, 0);
  for (int64_t c = 0; c < n_tiles; c += SCAN_THREADS) {
    int64_t i = c + threadIdx.x;
    int2 v = i < n_tiles ? tile_sums[i] : make_int2(0, 0);
    int2 total;
    int2 inc = block_scan_inclusive(v, &total);
    if (i < n_tiles) tile_sums[i] = make_int2(carry.x + inc.x - v.x, carry.y + inc.y - v.y);
    carry = add2(carry, total);
  }
}

// pass C: exclusive scans written out; element n holds the totals
__global__ void __launch_bounds__(SCAN_THREADS) graph_count_write(const int64_t* __restrict__ nbr,
                                                                  const int64_t* __restrict__ tgt_pos, int64_t n, int64_t k,
                                                                  int64_t n_datastore, int left_ctx, int right_ctx,
                                                                  int64_t invalid_ctx, const int2* __restrict__ tile_sums,
                                                                  int32_t* __restrict__ node_base,
                                                                  int32_t* __restrict__ valid_base) {
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int2 item[SCAN_ITEMS];
  int2 acc = make_int2(0, 0);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t idx = base + i;
    item[i] = make_int2(0, 0);
    if (idx < n) {
      int64_t pos = tgt_pos ? __ldg(tgt_pos + idx / k) : 0;
      ClusterShape s = cluster_shape(__ldg(nbr + idx), pos, n_datastore, left_ctx, right_ctx, invalid_ctx);
      item[i] = make_int2(s.valid ? 1 + s.nl + s.nr : 0, s.valid);
    }
    acc = add2(acc, item[i]);
  }
  int2 total;
  int2 inc = block_scan_inclusive(acc, &total);
  int2 run = add2(tile_sums[blockIdx.x], make_int2(inc.x - acc.x, inc.y - acc.y));
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t idx = base + i;
    if (idx < n) {
      node_base[idx] = run.x;
      valid_base[idx] = run.y;
    }
    run = add2(run, item[i]);
    if (idx == n - 1) {
      node_base[n] = run.x;
      valid_base[n] = run.y;
    }
  }
}

// sorted position -> node id inside a cluster (creation order: centre, left asc, right asc)
__device__ __forceinline__ int pos_to_id(int q, int nl) { return q == nl ? 0 : (q < nl ? q + 1 : q); }

__global__ void __launch_bounds__(256) graph_fill_kernel(const int64_t* __restrict__ nbr, const int64_t* __restrict__ tgt_pos,
                                                         int64_t n, int64_t k, int64_t n_datastore, int left_ctx,
                                                         int right_ctx, int64_t invalid_ctx,
                                                         const int32_t* __restrict__ node_base,
                                                         const int32_t* __restrict__ valid_base,
                                                         int64_t* __restrict__ ntgt_row, int32_t* __restrict__ ntgt_owner,
                                                         int32_t* __restrict__ ntgt_dist, int32_t* __restrict__ nn_indptr,
                                                         int32_t* __restrict__ nn_indices,
                                                         int32_t* __restrict__ inter_indptr,
                                                         int32_t* __restrict__ inter_indices,
                                                         int32_t* __restrict__ cluster_nl) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int64_t t = idx / k;
  const int base = node_base[idx];
  const int vb = valid_base[idx];
  if (inter_indptr) {
    if (idx % k == 0) inter_indptr[t] = vb;
    if (idx == n - 1) inter_indptr[t + 1] = valid_base[n];
  }
  if (idx == n - 1 && nn_indptr) {
    int nn = node_base[n], nv = valid_base[n];
    nn_indptr[nn] = 3 * nn - 2 * nv;
  }
  const int64_t o = __ldg(nbr + idx);
  const int64_t pos = tgt_pos ? __ldg(tgt_pos + t) : 0;
  const ClusterShape s = cluster_shape(o, pos, n_datastore, left_ctx, right_ctx, invalid_ctx);
  if (cluster_nl) cluster_nl[idx] = s.valid ? s.nl : -1;
  if (!s.valid) return;
  if (inter_indices) inter_indices[vb] = base;                 // centre node is created first (:367-374)
  const int nl = s.nl, nr = s.nr, w = 1 + nl + nr;
  const int e0 = 3 * base - 2 * vb;                            // every earlier cluster has 3w-2 edges
  const int deg0 = 1 + (nl > 0) + (nr > 0);
  for (int i = 0; i < w; ++i) {
    const int p = i == 0 ? nl : (i <= nl ? i - 1 : i);         // sorted position
    const int node = base + i;
    if (ntgt_row) ntgt_row[node] = o - nl + p;                 // rows are contiguous: o-nl .. o+nr
    if (ntgt_owner) ntgt_owner[node] = (int32_t)t;
    if (ntgt_dist) ntgt_dist[node] = p > nl ? p - nl : nl - p;
    if (nn_indptr) {
      int before;
      if (i == 0) before = 0;
      else if (i <= nl) before = deg0 + 2 * (i - 1) + (i >= 2 ? i - 2 : 0);
      else before = deg0 + (nl > 0 ? 2 * nl + (nl - 1) : 0) + 3 * (i - nl - 1);
      int e = e0 + before;
      nn_indptr[node] = e;
      if (nn_indices) {
        if (p > 0) nn_indices[e++] = base + pos_to_id(p - 1, nl);
        nn_indices[e++] = node;
        if (p < w - 1) nn_indices[e++] = base + pos_to_id(p + 1, nl);
      }
    }
  }
}

__global__ void __launch_bounds__(256) graph_tt_csr_kernel(int64_t B, int64_t L, int64_t intra_ctx,
                                                           int32_t* __restrict__ indptr, int32_t* __restrict__ indices) {
  // one thread per destination token; in-edges u in [max(0, v-ctx+1), v], ascending (row-major triu)
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= B * L) return;
  const int64_t b = g / L, v = g % L;
  const int64_t ctx = intra_ctx > 0 ? intra_ctx : L;
  // edges before v inside a block: sum_{j<v} min(j+1, ctx)
  auto before = [&](int64_t x) -> int64_t {
    if (x <= ctx) return x * (x + 1) / 2;
    return ctx * (ctx + 1) / 2 + (x - ctx) * ctx;
  };
  const int64_t per_block = before(L);
  int64_t e = b * per_block + before(v);
  indptr[g] = (int32_t)e;
  if (g == B * L - 1) indptr[B * L] = (int32_t)(B * per_block);
  const int64_t lo = v + 1 > ctx ? v + 1 - ctx : 0;
  for (int64_t u = lo; u <= v; ++u) indices[e++] = (int32_t)(b * L + u);
}

}  // namespace gnnlm

using namespace gnnlm;

extern "C" int64_t gnnlm_graph_workspace_bytes(int64_t n_clusters) {
  return (ceil_div(n_clusters, SCAN_TILE) + 1) * (int64_t)sizeof(int2);
}

extern "C" int32_t gnnlm_graph_count(const int64_t* nbr, const int64_t* tgt_pos, int64_t T, int64_t k,
                                     int64_t n_datastore, int32_t left_ctx, int32_t right_ctx, int64_t invalid_ctx,
                                     int32_t* node_base, int32_t* valid_base, void* workspace, int64_t workspace_bytes,
                                     gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(nbr && node_base && valid_base, GNNLM_E_ARG, "gnnlm_graph_count: null pointer");
  GNNLM_CHECK_ARG(T >= 0 && k > 0 && left_ctx >= 0 && right_ctx >= 0 && n_datastore > 0, GNNLM_E_SHAPE,
                  "gnnlm_graph_count: bad sizes T=%lld k=%lld", (long long)T, (long long)k);
  GNNLM_CHECK_ARG(invalid_ctx <= 0 || tgt_pos, GNNLM_E_ARG, "gnnlm_graph_count: tgt_pos required when invalid_ctx > 0");
  const int64_t n = T * k;
  GNNLM_CHECK_ARG(n * (1 + (int64_t)left_ctx + right_ctx) < (int64_t)INT32_MAX / 3, GNNLM_E_SHAPE,
                  "gnnlm_graph_count: node/edge ids overflow int32; split the batch");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    GNNLM_CUDA(cudaMemsetAsync(node_base, 0, sizeof(int32_t), st));
    GNNLM_CUDA(cudaMemsetAsync(valid_base, 0, sizeof(int32_t), st));
    return 0;
  }
  GNNLM_CHECK_ARG(workspace && workspace_bytes >= gnnlm_graph_workspace_bytes(n), GNNLM_E_WORKSPACE,
                  "gnnlm_graph_count: workspace too small");
  const int64_t tiles = ceil_div(n, SCAN_TILE);
  int2* sums = (int2*)workspace;
  graph_count_tiles<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(nbr, tgt_pos, n, k, n_datastore, left_ctx, right_ctx,
                                                               invalid_ctx, sums);
  graph_scan_tiles<<<1, SCAN_THREADS, 0, st>>>(sums, tiles);
  graph_count_write<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(nbr, tgt_pos, n, k, n_datastore, left_ctx, right_ctx,
                                                               invalid_ctx, sums, node_base, valid_base);
  GNNLM_LAUNCH_CHECK("gnnlm_graph_count");
  return 0;
}

extern "C" int32_t gnnlm_graph_fill(const int64_t* nbr, const int64_t* tgt_pos, int64_t T, int64_t k,
                                    int64_t n_datastore, int32_t left_ctx, int32_t right_ctx, int64_t invalid_ctx,
                                    const int32_t* node_base, const int32_t* valid_base, int64_t* ntgt_row,
                                    int32_t* ntgt_owner, int32_t* ntgt_dist, int32_t* nn_indptr, int32_t* nn_indices,
                                    int32_t* inter_indptr, int32_t* inter_indices, int32_t* cluster_nl,
                                    gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(nbr && node_base && valid_base, GNNLM_E_ARG, "gnnlm_graph_fill: null pointer");
  GNNLM_CHECK_ARG(!nn_indices || nn_indptr, GNNLM_E_ARG, "gnnlm_graph_fill: nn_indices needs nn_indptr");
  GNNLM_CHECK_ARG(invalid_ctx <= 0 || tgt_pos, GNNLM_E_ARG, "gnnlm_graph_fill: tgt_pos required when invalid_ctx > 0");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = T * k;
  if (n == 0) {
    if (nn_indptr) GNNLM_CUDA(cudaMemsetAsync(nn_indptr, 0, sizeof(int32_t), st));
    if (inter_indptr) GNNLM_CUDA(cudaMemsetAsync(inter_indptr, 0, sizeof(int32_t), st));
    return 0;
  }
  graph_fill_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(nbr, tgt_pos, n, k, n_datastore, left_ctx, right_ctx,
                                                                 invalid_ctx, node_base, valid_base, ntgt_row,
                                                                 ntgt_owner, ntgt_dist, nn_indptr, nn_indices,
                                                                 inter_indptr, inter_indices, cluster_nl);
  GNNLM_LAUNCH_CHECK("gnnlm_graph_fill");
  return 0;
}

extern "C" int64_t gnnlm_graph_tt_num_edges(int64_t B, int64_t L, int64_t intra_ctx) {
  const int64_t ctx = intra_ctx > 0 && intra_ctx < L ? intra_ctx : L;
  return B * (ctx * (ctx + 1) / 2 + (L - ctx) * ctx);
}

extern "C" int32_t gnnlm_graph_tt_csr(int64_t B, int64_t L, int64_t intra_ctx, int32_t* indptr, int32_t* indices,
                                      gnnlm_stream_t stream) {
  GNNLM_CHECK_ARG(indptr && indices, GNNLM_E_ARG, "gnnlm_graph_tt_csr: null pointer");
  GNNLM_CHECK_ARG(B > 0 && L > 0, GNNLM_E_SHAPE, "gnnlm_graph_tt_csr: bad sizes");
  GNNLM_CHECK_ARG(gnnlm_graph_tt_num_edges(B, L, intra_ctx) < INT32_MAX, GNNLM_E_SHAPE, "gnnlm_graph_tt_csr: too many edges");
  graph_tt_csr_kernel<<<(unsigned)ceil_div(B * L, 256), 256, 0, (cudaStream_t)stream>>>(B, L, intra_ctx, indptr, indices);
  GNNLM_LAUNCH_CHECK("gnnlm_graph_tt_csr");
  return 0;
}
