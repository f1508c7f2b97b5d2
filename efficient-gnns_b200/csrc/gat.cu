// Graph-attention aggregation (config 4): per-destination edge softmax + multi-head weighted neighbour sum.
//   DGL GATConv.forward  arxiv_dgl/models.py:196-217 : e = leaky_relu(el[src] + er[dst]); a = edge_softmax(e);
//                                                      out[dst,h,:] = sum_e a[e,h] * ft[src,h,:]
//   PyG GATConv (ppi_pyg/gnn.py:27-31) is the same computation with softmax_eps = 1e-16.
// CSR rows = destinations, col = sources, heads H, head width D (K = H*D floats per row).
// Forward : gat_edge_softmax (a[nnz,H]) + gat_aggregate (also used for d ft on the transposed graph through `eidx`).
// Backward: gat_bwd_rows (per destination: d a = <ft[src], d out[dst]> per head, softmax + leaky-relu backward,
//           d er, d pre[nnz,H]) + gat_segment_sum (d el over the transposed graph).
// Work split: one warp per chunk of rows (the SpMM chunk plan); rows above hub_threshold are taken by whole CTAs
// (8 warps stride over the edges, fixed-order shared-memory reduction) — no atomics, deterministic.
#include "common.cuh"

namespace b200gnn {

constexpr int GAT_THREADS = 256, GAT_WARPS = 8, GAT_MAXH = 16, GAT_MAXJ = 12;

__device__ __forceinline__ float gsum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
  return v;
}
__device__ __forceinline__ float gmax(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, d));
  return v;
}
__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : x * slope; }

// ---------------------------------------------------------------- edge softmax: a[e,h]
// one warp per destination row; lanes stride over the row's edges; three cheap passes over scalars.
__global__ void __launch_bounds__(GAT_THREADS) gat_edge_softmax_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ el,
    const float* __restrict__ er, int H, float slope, float eps, int64_t n_rows, float* __restrict__ a) {
  const int lane = threadIdx.x & 31;
  for (int64_t i = (int64_t)blockIdx.x * GAT_WARPS + (threadIdx.x >> 5); i < n_rows; i += (int64_t)gridDim.x * GAT_WARPS) {
    const int b = rowptr[i], e = rowptr[i + 1];
    if (b == e) continue;
    for (int h = 0; h < H; ++h) {
      const float r = er ? er[(size_t)i * H + h] : 0.f;
      float m = -INFINITY;
      for (int k = b + lane; k < e; k += 32) m = fmaxf(m, lrelu(el[(size_t)col[k] * H + h] + r, slope));
      m = gmax(m);
      float s = 0.f;
      for (int k = b + lane; k < e; k += 32) s += expf(lrelu(el[(size_t)col[k] * H + h] + r, slope) - m);
      s = gsum(s) + eps;
      for (int k = b + lane; k < e; k += 32)
        a[(size_t)k * H + h] = expf(lrelu(el[(size_t)col[k] * H + h] + r, slope) - m) / s;
    }
  }
}

struct GatAgg {
  const int32_t* rowptr; const int32_t* col; const int32_t* eidx;   // eidx: position of edge k in a[] (NULL: k)
  const int32_t* chunk_rowptr; const int32_t* hub_rows;
  const float* a; const float* ft; float* out;
  int64_t ldf, ldo;
  int32_t n_chunks, n_hub, hub_threshold, H, D, K;
};

// per-lane column map: lane handles vectors v = lane + 32*j (j < nj) of width W; head of vector v = (v*W)/D
template <typename V>
__device__ __forceinline__ void agg_edges(const GatAgg& p, int beg, int end, int stride, int first, int lane, int nj,
                                          V (&acc)[GAT_MAXJ]) {
  constexpr int W = VecTraits<V>::W;
  const V* F = reinterpret_cast<const V*>(p.ft);
  const size_t ldv = (size_t)(p.ldf / W);
  int head[GAT_MAXJ];
#pragma unroll
  for (int j = 0; j < GAT_MAXJ; ++j) head[j] = ((lane + 32 * j) * W) / p.D;
  for (int k = beg + first; k < end; k += stride) {
    const int src = __ldg(p.col + k);
    const size_t ak = (size_t)(p.eidx ? __ldg(p.eidx + k) : k) * p.H;
    const V* row = F + (size_t)src * ldv + lane;
#pragma unroll
    for (int j = 0; j < GAT_MAXJ; ++j)
      if (j < nj && (lane + 32 * j) * W < p.K) vfma(acc[j], __ldg(p.a + ak + head[j]), vldg(row + 32 * j));
  }
}

template <typename V>
__global__ void __launch_bounds__(GAT_THREADS) gat_aggregate_kernel(const GatAgg p) {
  constexpr int W = VecTraits<V>::W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = p.K / W, nj = (nvec + 31) / 32;
  V* O = reinterpret_cast<V*>(p.out);
  const size_t ldov = (size_t)(p.ldo / W);
  const int chunk = blockIdx.x * GAT_WARPS + warp;
  if (chunk >= p.n_chunks) return;
  const int r0 = __ldg(p.chunk_rowptr + chunk), r1 = __ldg(p.chunk_rowptr + chunk + 1);
  for (int r = r0; r < r1; ++r) {
    const int b = __ldg(p.rowptr + r), e = __ldg(p.rowptr + r + 1);
    if (e - b > p.hub_threshold) continue;
    V acc[GAT_MAXJ];
#pragma unroll
    for (int j = 0; j < GAT_MAXJ; ++j) vzero(acc[j]);
    agg_edges<V>(p, b, e, 1, 0, lane, nj, acc);
#pragma unroll
    for (int j = 0; j < GAT_MAXJ; ++j)
      if (j < nj && lane + 32 * j < nvec) O[(size_t)r * ldov + lane + 32 * j] = acc[j];
  }
}

// one CTA per hub row: warps stride over the edges, partials combined in warp order through shared memory
template <typename V>
__global__ void __launch_bounds__(GAT_THREADS) gat_aggregate_hub_kernel(const GatAgg p) {
  constexpr int W = VecTraits<V>::W;
  extern __shared__ float s_row[];   // K floats
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = p.K / W, nj = (nvec + 31) / 32;
  const int r = __ldg(p.hub_rows + blockIdx.x);
  const int b = __ldg(p.rowptr + r), e = __ldg(p.rowptr + r + 1);
  V acc[GAT_MAXJ];
#pragma unroll
  for (int j = 0; j < GAT_MAXJ; ++j) vzero(acc[j]);
  agg_edges<V>(p, b, e, GAT_WARPS, warp, lane, nj, acc);
  for (int i = threadIdx.x; i < p.K; i += GAT_THREADS) s_row[i] = 0.f;
  __syncthreads();
  V* sv = reinterpret_cast<V*>(s_row);
  for (int w = 0; w < GAT_WARPS; ++w) {
    if (warp == w) {
#pragma unroll
      for (int j = 0; j < GAT_MAXJ; ++j)
        if (j < nj && lane + 32 * j < nvec) { V t = sv[lane + 32 * j]; vadd(t, acc[j]); sv[lane + 32 * j] = t; }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < p.K; i += GAT_THREADS) p.out[(size_t)r * p.ldo + i] = s_row[i];
}

// ---------------------------------------------------------------- backward, per destination row
struct GatBwd {
  const int32_t* rowptr; const int32_t* col; const float* a; const float* ft; const float* dout;
  const float* el; const float* er;
  float* dpre;   // [nnz,H]  out: d loss / d (el[src]+er[dst])
  float* der;    // [n_rows,H] out (may be NULL when there is no er)
  const int32_t* row_list;       // CTA kernel: rows to process (NULL: all rows)
  const int32_t* chunk_rowptr;   // warp kernel: chunk plan
  int64_t ldf, ldd, n_rows;
  int32_t n_list, n_chunks, hub_threshold;
  int32_t H, D, K;
  float slope;
};

// One CTA per destination row (rows are cheap on average; a CTA gives hubs 8 warps).  Phase 1: d a[e,h] =
// <ft[src,h,:], dout[dst,h,:]> (warp per edge, per-head warp reductions), phase 2: softmax and leaky-relu backward.
__global__ void __launch_bounds__(GAT_THREADS) gat_bwd_rows_kernel(const GatBwd p) {
  __shared__ float s_S[GAT_WARPS][GAT_MAXH];
  __shared__ float s_tot[GAT_MAXH];
  __shared__ float s_der[GAT_WARPS][GAT_MAXH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_items = p.row_list ? p.n_list : p.n_rows;
  for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
    const int64_t i = p.row_list ? p.row_list[it] : it;
    const int b = p.rowptr[i], e = p.rowptr[i + 1];
    float Sh[GAT_MAXH];
#pragma unroll
    for (int h = 0; h < GAT_MAXH; ++h) Sh[h] = 0.f;
    const float* g = p.dout + (size_t)i * p.ldd;
    for (int k = b + warp; k < e; k += GAT_WARPS) {
      const float* f = p.ft + (size_t)p.col[k] * p.ldf;
      float part[GAT_MAXH];
#pragma unroll
      for (int h = 0; h < GAT_MAXH; ++h) part[h] = 0.f;
      for (int c = lane; c < p.K; c += 32) {
        const float v = __ldg(f + c) * __ldg(g + c);
        const int h = c / p.D;
#pragma unroll
        for (int hh = 0; hh < GAT_MAXH; ++hh) part[hh] += (hh == h) ? v : 0.f;
      }
#pragma unroll
      for (int h = 0; h < GAT_MAXH; ++h)
        if (h < p.H) {
          const float da = gsum(part[h]);
          if (lane == 0) p.dpre[(size_t)k * p.H + h] = da;          // staged: d a
          Sh[h] += p.a[(size_t)k * p.H + h] * da;                   // same value on every lane
        }
    }
    if (lane == 0)
      for (int h = 0; h < p.H; ++h) s_S[warp][h] = Sh[h];
    __syncthreads();
    if (threadIdx.x < p.H) {
      float t = 0.f;
      for (int w = 0; w < GAT_WARPS; ++w) t += s_S[w][threadIdx.x];
      s_tot[threadIdx.x] = t;
    }
    __syncthreads();
    // phase 2: d e = a (d a - S);  d pre = d e * leaky'(pre);  d er[i,h] = sum_e d pre
    float dr[GAT_MAXH];
#pragma unroll
    for (int h = 0; h < GAT_MAXH; ++h) dr[h] = 0.f;
    for (int k = b + threadIdx.x; k < e; k += GAT_THREADS) {
      const int src = p.col[k];
      for (int h = 0; h < p.H; ++h) {
        const size_t o = (size_t)k * p.H + h;
        const float de = p.a[o] * (p.dpre[o] - s_tot[h]);
        const float pre = p.el[(size_t)src * p.H + h] + (p.er ? p.er[(size_t)i * p.H + h] : 0.f);
        const float dp = pre > 0.f ? de : de * p.slope;
        p.dpre[o] = dp;
        dr[h] += dp;
      }
    }
    if (p.der) {
      for (int h = 0; h < p.H; ++h) {
        const float t = gsum(dr[h]);
        if (lane == 0) s_der[warp][h] = t;
      }
      __syncthreads();
      if (threadIdx.x < p.H) {
        float t = 0.f;
        for (int w = 0; w < GAT_WARPS; ++w) t += s_der[w][threadIdx.x];
        p.der[(size_t)i * p.H + threadIdx.x] = t;
      }
    }
    __syncthreads();
  }
}


// Fast path: D % 4 == 0 and D/4 (lanes per head) a power of two <= 32 (D in {4,8,...,128}).  One warp per chunk of rows; per edge the
// per-head dot product is a segmented xor-shuffle reduction over the D/4 lanes that hold that head; hub rows are
// left to the CTA kernel above.
constexpr int GAT_FJ = 8;   // float4 vectors per lane: K <= 1024
__global__ void __launch_bounds__(GAT_THREADS) gat_bwd_rows_warp_kernel(const GatBwd p) {
  __shared__ float s_S[GAT_WARPS][GAT_MAXH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = blockIdx.x * GAT_WARPS + warp;
  if (chunk >= p.n_chunks) return;
  const int lph = p.D >> 2;                      // lanes per head
  const int nvec = p.K >> 2, nj = (nvec + 31) >> 5;
  const float4* F = reinterpret_cast<const float4*>(p.ft);
  const float4* G = reinterpret_cast<const float4*>(p.dout);
  const size_t ldfv = (size_t)(p.ldf >> 2), lddv = (size_t)(p.ldd >> 2);
  int head[GAT_FJ];
#pragma unroll
  for (int j = 0; j < GAT_FJ; ++j) head[j] = (lane + 32 * j) / lph;
  const int r0 = __ldg(p.chunk_rowptr + chunk), r1 = __ldg(p.chunk_rowptr + chunk + 1);
  for (int i = r0; i < r1; ++i) {
    const int b = __ldg(p.rowptr + i), e = __ldg(p.rowptr + i + 1);
    if (e - b > p.hub_threshold || b == e) {
      if (b == e && p.der && lane < p.H) p.der[(size_t)i * p.H + lane] = 0.f;
      continue;
    }
    float4 g[GAT_FJ];
    float S[GAT_FJ];
#pragma unroll
    for (int j = 0; j < GAT_FJ; ++j) {
      S[j] = 0.f;
      g[j] = (j < nj && lane + 32 * j < nvec) ? __ldg(G + (size_t)i * lddv + lane + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int k = b; k < e; ++k) {
      const float4* f = F + (size_t)__ldg(p.col + k) * ldfv + lane;
#pragma unroll
      for (int j = 0; j < GAT_FJ; ++j) {
        if (j < nj) {
          float d = 0.f;
          if (lane + 32 * j < nvec) { const float4 x = __ldg(f + 32 * j); d = x.x * g[j].x + x.y * g[j].y + x.z * g[j].z + x.w * g[j].w; }
          for (int o = 1; o < lph; o <<= 1) d += __shfl_xor_sync(FULL_MASK, d, o);   // lph <= 32: stays inside the head's lanes
          if (lane + 32 * j < nvec) {
            const size_t o = (size_t)k * p.H + head[j];
            if ((lane % lph) == 0) p.dpre[o] = d;
            S[j] = fmaf(__ldg(p.a + o), d, S[j]);
          }
        }
      }
    }
    // per-head totals: each head lives in one aligned group of lph lanes of one chunk j; its leader publishes S
    __syncwarp();
#pragma unroll
    for (int j = 0; j < GAT_FJ; ++j)
      if (j < nj && lane + 32 * j < nvec && (lane % lph) == 0) s_S[warp][head[j]] = S[j];
    __syncwarp();
    float dr[GAT_MAXH];
#pragma unroll
    for (int h = 0; h < GAT_MAXH; ++h) dr[h] = 0.f;
    for (int k = b + lane; k < e; k += 32) {
      const int src = __ldg(p.col + k);
#pragma unroll
      for (int h = 0; h < GAT_MAXH; ++h)
        if (h < p.H) {
          const size_t o = (size_t)k * p.H + h;
          const float de = p.a[o] * (p.dpre[o] - s_S[warp][h]);
          const float pre = p.el[(size_t)src * p.H + h] + (p.er ? p.er[(size_t)i * p.H + h] : 0.f);
          const float dp = pre > 0.f ? de : de * p.slope;
          p.dpre[o] = dp;
          dr[h] += dp;
        }
    }
    if (p.der) {
#pragma unroll
      for (int h = 0; h < GAT_MAXH; ++h)
        if (h < p.H) { const float t = gsum(dr[h]); if (lane == 0) p.der[(size_t)i * p.H + h] = t; }
    }
    __syncwarp();
  }
}

// out[j,h] = sum over rows' edges k of vals[eidx[k], h]   (d el over the transposed graph)
__global__ void __launch_bounds__(GAT_THREADS) gat_segment_sum_kernel(const int32_t* __restrict__ rowptr,
                                                                      const int32_t* __restrict__ eidx, const float* __restrict__ vals,
                                                                      int H, int64_t n_rows, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (int64_t j = (int64_t)blockIdx.x * GAT_WARPS + (threadIdx.x >> 5); j < n_rows; j += (int64_t)gridDim.x * GAT_WARPS) {
    const int b = rowptr[j], e = rowptr[j + 1];
    for (int h = 0; h < H; ++h) {
      float s = 0.f;
      for (int k = b + lane; k < e; k += 32) s += vals[(size_t)(eidx ? eidx[k] : k) * H + h];
      s = gsum(s);
      if (lane == 0) out[(size_t)j * H + h] = s;
    }
  }
}

static inline int rows_grid(int64_t n) {
  int64_t g = (n + GAT_WARPS - 1) / GAT_WARPS;
  if (g > 148 * 16) g = 148 * 16;
  return (int)(g < 1 ? 1 : g);
}

template <typename V>
static int launch_agg(const GatAgg& p, cudaStream_t st) {
  int rc;
  gat_aggregate_kernel<V><<<(p.n_chunks + GAT_WARPS - 1) / GAT_WARPS, GAT_THREADS, 0, st>>>(p);
  if ((rc = check_launch())) return rc;
  if (p.n_hub > 0) {
    gat_aggregate_hub_kernel<V><<<p.n_hub, GAT_THREADS, p.K * sizeof(float), st>>>(p);
    if ((rc = check_launch())) return rc;
  }
  return B200GNN_OK;
}

}  // namespace b200gnn

using namespace b200gnn;

extern "C" int b200gnn_gat_edge_softmax_f32(const int32_t* rowptr, const int32_t* col, const float* el, const float* er,
                                            int64_t n_rows, int64_t H, float negative_slope, float softmax_eps, float* a,
                                            void* stream) {
  if (!rowptr || !el || !a || n_rows < 0 || H <= 0 || H > GAT_MAXH) return B200GNN_ERR_BAD_ARG;
  if (n_rows == 0) return B200GNN_OK;
  gat_edge_softmax_kernel<<<rows_grid(n_rows), GAT_THREADS, 0, (cudaStream_t)stream>>>(rowptr, col, el, er, (int)H,
                                                                                      negative_slope, softmax_eps, n_rows, a);
  return check_launch();
}

extern "C" int b200gnn_gat_aggregate_f32(const int32_t* rowptr, const int32_t* col, const int32_t* eidx, const float* a,
                                         const float* ft, int64_t ldf, float* out, int64_t ldo, int64_t n_rows, int64_t H,
                                         int64_t D, const int32_t* chunk_rowptr, int64_t n_chunks, int32_t hub_threshold,
                                         const int32_t* hub_rows, int64_t n_hub, void* stream) {
  const int64_t K = H * D;
  if (!rowptr || !a || !ft || !out || n_rows < 0 || H <= 0 || D <= 0 || H > GAT_MAXH || ldf < K || ldo < K || !chunk_rowptr ||
      n_chunks < 0 || n_hub < 0 || (n_hub > 0 && !hub_rows))
    return B200GNN_ERR_BAD_ARG;
  if (n_rows == 0 || n_chunks == 0) return B200GNN_OK;
  GatAgg p;
  p.rowptr = rowptr; p.col = col; p.eidx = eidx; p.chunk_rowptr = chunk_rowptr; p.hub_rows = hub_rows;
  p.a = a; p.ft = ft; p.out = out; p.ldf = ldf; p.ldo = ldo;
  p.n_chunks = (int32_t)n_chunks; p.n_hub = (int32_t)n_hub; p.hub_threshold = hub_threshold;
  p.H = (int32_t)H; p.D = (int32_t)D; p.K = (int32_t)K;
  cudaStream_t st = (cudaStream_t)stream;
  // a vector must not straddle two heads: D % W == 0
  if (D % 4 == 0 && ldf % 4 == 0 && ldo % 4 == 0 && aligned_to(ft, 16) && aligned_to(out, 16) && K <= 4 * 32 * GAT_MAXJ)
    return launch_agg<float4>(p, st);
  if (D % 2 == 0 && ldf % 2 == 0 && ldo % 2 == 0 && aligned_to(ft, 8) && aligned_to(out, 8) && K <= 2 * 32 * GAT_MAXJ)
    return launch_agg<float2>(p, st);
  if (K <= 32 * GAT_MAXJ) return launch_agg<float>(p, st);
  return B200GNN_ERR_UNSUPPORTED;
}

extern "C" int b200gnn_gat_bwd_rows_f32(const int32_t* rowptr, const int32_t* col, const float* a, const float* ft, int64_t ldf,
                                        const float* dout, int64_t ldd, const float* el, const float* er, int64_t n_rows,
                                        int64_t H, int64_t D, float negative_slope, float* dpre, float* der,
                                        const int32_t* chunk_rowptr, int64_t n_chunks, int32_t hub_threshold,
                                        const int32_t* hub_rows, int64_t n_hub, void* stream) {
  if (!rowptr || !a || !ft || !dout || !el || !dpre || n_rows < 0 || H <= 0 || D <= 0 || H > GAT_MAXH || ldf < H * D ||
      ldd < H * D)
    return B200GNN_ERR_BAD_ARG;
  if (n_rows == 0) return B200GNN_OK;
  GatBwd p;
  p.rowptr = rowptr; p.col = col; p.a = a; p.ft = ft; p.dout = dout; p.el = el; p.er = er; p.dpre = dpre; p.der = der;
  p.ldf = ldf; p.ldd = ldd; p.n_rows = n_rows; p.H = (int32_t)H; p.D = (int32_t)D; p.K = (int32_t)(H * D);
  p.slope = negative_slope;
  p.row_list = nullptr; p.chunk_rowptr = chunk_rowptr; p.n_list = 0; p.n_chunks = (int32_t)n_chunks;
  p.hub_threshold = hub_threshold;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t lph = D / 4;
  const bool fast = chunk_rowptr && n_chunks > 0 && D % 4 == 0 && (lph & (lph - 1)) == 0 && H * D <= 4 * 32 * GAT_FJ &&
                    ldf % 4 == 0 && ldd % 4 == 0 && aligned_to(ft, 16) && aligned_to(dout, 16) && (n_hub == 0 || hub_rows) &&
                    lph <= 32;
  int rc;
  if (fast) {
    gat_bwd_rows_warp_kernel<<<(int)((n_chunks + GAT_WARPS - 1) / GAT_WARPS), GAT_THREADS, 0, st>>>(p);
    if ((rc = check_launch())) return rc;
    if (n_hub > 0) {
      p.row_list = hub_rows; p.n_list = (int32_t)n_hub;
      gat_bwd_rows_kernel<<<(int)n_hub, GAT_THREADS, 0, st>>>(p);
      if ((rc = check_launch())) return rc;
    }
    return B200GNN_OK;
  }
  int64_t grid = n_rows < 148 * 32 ? n_rows : 148 * 32;
  gat_bwd_rows_kernel<<<(int)grid, GAT_THREADS, 0, st>>>(p);
  return check_launch();
}

extern "C" int b200gnn_segment_sum_heads_f32(const int32_t* rowptr, const int32_t* eidx, const float* vals, int64_t n_rows,
                                             int64_t H, float* out, void* stream) {
  if (!rowptr || !vals || !out || n_rows < 0 || H <= 0) return B200GNN_ERR_BAD_ARG;
  if (n_rows == 0) return B200GNN_OK;
  gat_segment_sum_kernel<<<rows_grid(n_rows), GAT_THREADS, 0, (cudaStream_t)stream>>>(rowptr, eidx, vals, (int)H, n_rows, out);
  return check_launch();
}
