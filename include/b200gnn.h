/*
 * b200gnn.h — C ABI of the B200-native sparse message-passing engine.
 *
 * This is the drop-in boundary for the hot path of chaitjo/efficient-gnns
 * (SURVEY.md §8b).  The reference reaches its sparse arithmetic through
 * un-vendored Python/C++ dependencies (torch_sparse / torch_scatter / PyG);
 * each entry point below names the reference call site (file:line, relative
 * to /root/reference) whose arithmetic it replaces.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the name ends in `_host`.
 *   - The library never allocates or frees: callers own every buffer,
 *     including workspaces (size helpers are provided).
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *     performs no host<->device synchronisation and is safe under CUDA-graph
 *     capture.
 *   - Return value: 0 on success, a negative B200GNN_ERR_* otherwise.
 *     b200gnn_last_cuda_error() gives the CUDA error string of the last
 *     B200GNN_ERR_CUDA on the calling thread.
 *   - Engine-side indices are int32 (the Python host narrows the reference's
 *     int64 once, with a range check); features are fp32 row-major with an
 *     explicit leading dimension (in elements).
 */
#ifndef B200GNN_H_
#define B200GNN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200GNN_OK 0
#define B200GNN_ERR_BAD_ARG (-1)
#define B200GNN_ERR_UNSUPPORTED (-2)
#define B200GNN_ERR_CUDA (-3)

#define B200GNN_REDUCE_SUM 0
#define B200GNN_REDUCE_MEAN 1

#define B200GNN_ABI_VERSION 1

int b200gnn_abi_version(void);
const char* b200gnn_error_string(int code);
const char* b200gnn_last_cuda_error(void);
/* Number of kernel launches issued through this library by the calling
 * process since load / since the last reset (bench.py's gpu_launches). */
int64_t b200gnn_launch_count(void);
void b200gnn_reset_launch_count(void);

/* ------------------------------------------------------------------ *
 * CSR hub plan.  Rows whose degree exceeds `hub_threshold` are split into
 * segments of `seg_len` non-zeros processed by whole CTAs, so one hub node
 * (ARXIV-shape: degree ~2e4) cannot serialise a warp.  Built once per graph
 * and cached next to rowptr, like torch_sparse's SparseStorage caches
 * rowcount/colptr/csr2csc (used via arxiv_pyg/gnn.py:236-240).
 *   counts_out: int32[2] = {n_hub_rows, n_segments}
 *   hub_rows:   int32[n_hub]  ascending row ids
 *   hub_segptr: int32[n_hub+1] exclusive prefix of per-row segment counts
 * ------------------------------------------------------------------ */
int b200gnn_csr_hub_count(const int32_t* rowptr, int64_t n_rows,
                          int32_t hub_threshold, int32_t seg_len,
                          int32_t* counts_out, void* stream);
int b200gnn_csr_hub_fill(const int32_t* rowptr, int64_t n_rows,
                         int32_t hub_threshold, int32_t seg_len,
                         int32_t* hub_rows, int32_t* hub_segptr,
                         int64_t n_hub, void* stream);

/* ------------------------------------------------------------------ *
 * Row-segmented CSR SpMM  Y[i,:] = reduce_{e in row i} val[e] * X[col[e],:]
 * Replaces torch_sparse spmm_sum / spmm_mean reached from
 *   GCNConv  arxiv_pyg/gnn.py:47,52   (reduce=sum, weighted)
 *   SAGEConv arxiv_pyg/gnn.py:79,84   (reduce=mean, val==NULL)
 *   adj_t.matmul(x, reduce='mean')  mag_pyg/gnn.py:162
 * and their backward (the same kernel on the CSC view).
 *   val  : NULL => all ones.
 *   bias : NULL or float[K]; added after the reduction (GCNConv `out += bias`).
 *   stat_partial : NULL or float[b200gnn_spmm_stat_slots()][2][K]; every slot
 *          receives a deterministic partial column sum (slot,0,:) and sum of
 *          squares (slot,1,:) of the rows of Y it produced, for the
 *          BatchNorm1d that follows the conv (arxiv_pyg/gnn.py:48).
 *   hub_* : plan from b200gnn_csr_hub_*; n_hub==0 disables the split path
 *          (then hub_threshold must be >= the maximum degree or INT32_MAX).
 *   hub_workspace : float[n_seg][K] scratch for segment partials.
 * MEAN divides by max(degree,1); empty rows give 0 (+bias).
 * ------------------------------------------------------------------ */
int64_t b200gnn_spmm_stat_slots(int64_t n_rows, int64_t n_hub);
int b200gnn_spmm_csr_f32(const int32_t* rowptr, const int32_t* col,
                         const float* val, const float* X, int64_t ldx,
                         float* Y, int64_t ldy, int64_t n_rows, int64_t n_src,
                         int64_t K, int reduce, const float* bias,
                         float* stat_partial, int32_t hub_threshold,
                         int32_t seg_len, const int32_t* hub_rows,
                         const int32_t* hub_segptr, int64_t n_hub,
                         int64_t n_seg, float* hub_workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200GNN_H_ */
