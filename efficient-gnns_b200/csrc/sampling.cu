// Mini-batch graph sampling on the device: what the reference's GraphSAINT loader does on CPU workers with
// torch_sparse (`GraphSAINTRandomWalkSampler(homo_data, batch_size, walk_length=num_layers, num_steps, sample_coverage=0)`,
// mag_pyg/gnn.py:361-366 — roots uniform over the nodes, one uniform random walk per root, the induced subgraph of the
// visited nodes, every node / edge attribute sliced along).
//
//   random_walk     one thread per walker; step s of walker w draws word (s % 4) of Philox4x32-10(seed, offset, w * ceil(L/4)
//                   + s / 4) and moves to col[rowptr[v] + (r * deg >> 32)] — a uniform neighbour; a node without
//                   out-edges holds the walker (torch_sparse.random_walk's rule).  A pure function of (seed, offset, w), so
//                   the oracle restates it bit for bit.  The graph (4 B / edge) is L2-resident after the first step.
//   saint_subgraph  induced subgraph of a SORTED UNIQUE node set S over CSR: a node -> local-id map, one warp per selected
//                   row counting / writing the edges whose column is in S (ballot prefix: CSR order preserved, as
//                   SparseTensor.saint_subgraph keeps it), with the edge ids of the parent graph for attribute slicing.
// Integer / index work: bit-exact against oracle/sampling.py.
#include "common.cuh"
#include "philox.cuh"

namespace b200gnn {
namespace sampling {

__global__ void __launch_bounds__(256) random_walk_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                          const int64_t* __restrict__ start, int64_t n_walks, int walk_length,
                                                          uint64_t seed, uint64_t offset, int64_t* __restrict__ out) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_walks) return;
  const int blocks_per_walk = (walk_length + 3) / 4;
  int64_t v = start[w];
  int64_t* o = out + w * (walk_length + 1);
  o[0] = v;
  uint4 r = make_uint4(0, 0, 0, 0);
  for (int s = 0; s < walk_length; ++s) {
    if ((s & 3) == 0) r = philox4x32(seed, offset, (uint64_t)w * blocks_per_walk + (s >> 2));
    const uint32_t u = (s & 3) == 0 ? r.x : (s & 3) == 1 ? r.y : (s & 3) == 2 ? r.z : r.w;
    const int32_t b = __ldg(rowptr + v), e = __ldg(rowptr + v + 1);
    const uint32_t deg = (uint32_t)(e - b);
    if (deg > 0) v = __ldg(col + b + (int32_t)(((uint64_t)u * deg) >> 32));
    o[s + 1] = v;
  }
}

__global__ void __launch_bounds__(256) fill_map_kernel(const int64_t* __restrict__ nodes, int64_t n_sel, int32_t* __restrict__ map) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_sel) map[nodes[i]] = (int32_t)i;
}

// FILL = false: counts[i] = kept edges of selected row i.  FILL = true: writes them at out_ptr[i]...
template <bool FILL>
__global__ void __launch_bounds__(256) induced_rows_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                           const int64_t* __restrict__ eid, const int64_t* __restrict__ nodes,
                                                           int64_t n_sel, const int32_t* __restrict__ map,
                                                           int64_t* __restrict__ counts_or_ptr, int64_t* __restrict__ out_row,
                                                           int64_t* __restrict__ out_col, int64_t* __restrict__ out_eid) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n_sel) return;
  const int64_t v = nodes[i];
  const int32_t b = rowptr[v], e = rowptr[v + 1];
  int64_t base = FILL ? counts_or_ptr[i] : 0;
  int64_t cnt = 0;
  for (int32_t j = b + lane; j < ((e - b + 31) / 32) * 32 + b; j += 32) {
    const int32_t c = j < e ? map[col[j]] : -1;
    const unsigned m = __ballot_sync(FULL_MASK, c >= 0);
    if (FILL && c >= 0) {
      const int64_t o = base + __popc(m & ((1u << lane) - 1));
      out_row[o] = i;
      out_col[o] = c;
      out_eid[o] = eid ? eid[j] : (int64_t)j;
    }
    base += __popc(m);
    cnt += __popc(m);
  }
  if (!FILL && lane == 0) counts_or_ptr[i] = cnt;
}

}  // namespace sampling
}  // namespace b200gnn

using namespace b200gnn;

extern "C" int b200gnn_random_walk_i64(const int32_t* rowptr, const int32_t* col, int64_t n_nodes, const int64_t* start,
                                       int64_t n_walks, int32_t walk_length, uint64_t seed, uint64_t offset, int64_t* out,
                                       void* stream) {
  if (!rowptr || !start || !out || n_nodes <= 0 || n_walks < 0 || walk_length < 0 || walk_length > 4096) return B200GNN_ERR_BAD_ARG;
  if (n_walks == 0) return B200GNN_OK;
  sampling::random_walk_kernel<<<(unsigned)((n_walks + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      rowptr, col, start, n_walks, walk_length, seed, offset, out);
  return check_launch();
}

// node_map: int32 [n_nodes] workspace holding -1 everywhere on entry; the call sets map[v] = position of v in `nodes` for the
// selected nodes.  The caller restores those entries to -1 after the fill call (so one map serves every batch without an
// O(N) memset per batch).
extern "C" int b200gnn_saint_subgraph_count_i64(const int32_t* rowptr, const int32_t* col, const int64_t* nodes, int64_t n_sel,
                                                int32_t* node_map, int64_t* counts, void* stream) {
  if (!rowptr || !nodes || !node_map || !counts || n_sel < 0) return B200GNN_ERR_BAD_ARG;
  if (n_sel == 0) return B200GNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  sampling::fill_map_kernel<<<(unsigned)((n_sel + 255) / 256), 256, 0, st>>>(nodes, n_sel, node_map);
  int rc;
  if ((rc = check_launch())) return rc;
  sampling::induced_rows_kernel<false><<<(unsigned)((n_sel * 32 + 255) / 256), 256, 0, st>>>(rowptr, col, nullptr, nodes, n_sel, node_map,
                                                                                        counts, nullptr, nullptr, nullptr);
  return check_launch();
}

// out_ptr: exclusive prefix sums of `counts` (int64 [n_sel]); eid: optional parent edge ids per CSR position (NULL: the CSR
// position itself).  Outputs: local row, local column, parent edge id, CSR order.
extern "C" int b200gnn_saint_subgraph_fill_i64(const int32_t* rowptr, const int32_t* col, const int64_t* eid, const int64_t* nodes,
                                               int64_t n_sel, const int32_t* node_map, const int64_t* out_ptr, int64_t* out_row,
                                               int64_t* out_col, int64_t* out_eid, void* stream) {
  if (!rowptr || !nodes || !node_map || !out_ptr || !out_row || !out_col || !out_eid || n_sel < 0) return B200GNN_ERR_BAD_ARG;
  if (n_sel == 0) return B200GNN_OK;
  sampling::induced_rows_kernel<true><<<(unsigned)((n_sel * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      rowptr, col, eid, nodes, n_sel, node_map, const_cast<int64_t*>(out_ptr), out_row, out_col, out_eid);
  return check_launch();
}
