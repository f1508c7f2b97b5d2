// Graph-attention aggregation (config 4): per-destination edge softmax + multi-head weighted neighbour sum.
//   DGL GATConv.forward  arxiv_dgl/models.py:196-217 : e = leaky_relu(el[src] + er[dst]); a = edge_softmax(e);
//                                                      out[dst,h,:] = sum_e a[e,h] * ft[src,h,:]
//   PyG GATConv (ppi_pyg/gnn.py:27-31) is the same computation with softmax_eps = 1e-16.
// CSR rows = destinations, col = sources, heads H, head width D (K = H*D floats per row).
// Forward : gat_edge_softmax (a[nnz,H]) + gat_aggregate (also used for d ft on the transposed graph through `eidx`).
// Backward: gat_bwd_rows (per destination: d a = <ft[src], d out[dst]> per head, softmax + leaky-relu backward,
//           d er, d pre[nnz,H]) + gat_segment_sum (d el over the transposed graph).
// Work split.  Row-wide kernels (aggregate, bwd_rows): one warp per chunk of rows (the SpMM chunk plan); lanes own
// vectors lane+32j of the K-float row and U edges are loaded before they are consumed, so U*NJ independent row
// gathers are in flight per warp.  Rows above hub_threshold are split into seg_len-edge segments (the SpMM hub plan):
// one CTA per segment, scheduled as the first CTAs of the same launch, 8 warps striding over groups of U edges and
// combining in warp order through shared memory into a per-segment partial; a small finalize kernel adds a hub
// row's segments in order (and, in the backward, runs the softmax pass once S is complete).  Scalar kernels (edge softmax,
// segment sum): lanes stride over a row's edges and carry all H heads at once; rows above GAT_CTA_DEG are taken
// by the whole CTA.  No atomics anywhere: every sum has a fixed order.
#include "common.cuh"

namespace b200gnn {

constexpr int GAT_THREADS = 256, GAT_WARPS = 8, GAT_MAXH = 16, GAT_MAXJ = 12;
constexpr int GAT_CTA_DEG = 512;

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
__device__ __forceinline__ float vdot(const float& a, const float& b) { return a * b; }
__device__ __forceinline__ float vdot(const float2& a, const float2& b) { return fmaf(a.x, b.x, a.y * b.y); }
__device__ __forceinline__ float vdot(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// per-head reduction over the group that owns a row: a warp (shuffles) or the 8-warp CTA (shuffles + shared memory,
// combined in warp order).  Every thread of the group gets the result.
template <bool CTA, bool MAX>
__device__ __forceinline__ void reduce_heads(float (&v)[GAT_MAXH], int H, float (*s_red)[GAT_MAXH], int lane, int warp) {
#pragma unroll
  for (int h = 0; h < GAT_MAXH; ++h)
    if (h < H) v[h] = MAX ? gmax(v[h]) : gsum(v[h]);
  if (CTA) {
    if (lane == 0) {
#pragma unroll
      for (int h = 0; h < GAT_MAXH; ++h)
        if (h < H) s_red[warp][h] = v[h];
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < GAT_MAXH; ++h)
      if (h < H) {
        float t = s_red[0][h];
        for (int w = 1; w < GAT_WARPS; ++w) t = MAX ? fmaxf(t, s_red[w][h]) : t + s_red[w][h];
        v[h] = t;
      }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- edge softmax: a[e,h]
struct GatSoftmax {
  const int32_t* rowptr; const int32_t* col; const float* el; const float* er; float* a;
  const uint8_t* keep;   // [nnz] or NULL: edges with keep == 0 are dropped (a = 0, excluded from the softmax) — edge_drop
  int64_t n_rows; int32_t H; float slope, eps;
};

template <bool CTA>
__device__ __forceinline__ void softmax_row(const GatSoftmax& p, int64_t i, int b, int e, int tid, int nt,
                                            float (*s_red)[GAT_MAXH], int lane, int warp) {
  float r[GAT_MAXH], m[GAT_MAXH], s[GAT_MAXH];
#pragma unroll
  for (int h = 0; h < GAT_MAXH; ++h) {
    r[h] = (p.er && h < p.H) ? __ldg(p.er + (size_t)i * p.H + h) : 0.f;
    m[h] = -INFINITY;
    s[h] = 0.f;
  }
  for (int k = b + tid; k < e; k += nt) {
    if (p.keep && !p.keep[k]) continue;
    const float* x = p.el + (size_t)__ldg(p.col + k) * p.H;
#pragma unroll
    for (int h = 0; h < GAT_MAXH; ++h)
      if (h < p.H) m[h] = fmaxf(m[h], lrelu(__ldg(x + h) + r[h], p.slope));
  }
  reduce_heads<CTA, true>(m, p.H, s_red, lane, warp);
  for (int k = b + tid; k < e; k += nt) {
    if (p.keep && !p.keep[k]) continue;
    const float* x = p.el + (size_t)__ldg(p.col + k) * p.H;
#pragma unroll
    for (int h = 0; h < GAT_MAXH; ++h)
      if (h < p.H) s[h] += expf(lrelu(__ldg(x + h) + r[h], p.slope) - m[h]);
  }
  reduce_heads<CTA, false>(s, p.H, s_red, lane, warp);
  for (int k = b + tid; k < e; k += nt) {
    const bool kept = !p.keep || p.keep[k];
    const float* x = p.el + (size_t)__ldg(p.col + k) * p.H;
#pragma unroll
    for (int h = 0; h < GAT_MAXH; ++h)
      if (h < p.H) p.a[(size_t)k * p.H + h] = kept ? expf(lrelu(__ldg(x + h) + r[h], p.slope) - m[h]) / (s[h] + p.eps) : 0.f;
  }
}

__global__ void __launch_bounds__(GAT_THREADS) gat_edge_softmax_kernel(const GatSoftmax p) {
  __shared__ float s_red[GAT_WARPS][GAT_MAXH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t r0 = (int64_t)blockIdx.x * GAT_WARPS; r0 < p.n_rows; r0 += (int64_t)gridDim.x * GAT_WARPS) {
    const int64_t i = r0 + warp;
    if (i < p.n_rows) {
      const int b = __ldg(p.rowptr + i), e = __ldg(p.rowptr + i + 1);
      if (e > b && e - b <= GAT_CTA_DEG) softmax_row<false>(p, i, b, e, lane, 32, s_red, lane, warp);
    }
    for (int w = 0; w < GAT_WARPS; ++w) {          // same decision in every thread of the CTA
      const int64_t r = r0 + w;
      if (r >= p.n_rows) break;
      const int b = __ldg(p.rowptr + r), e = __ldg(p.rowptr + r + 1);
      if (e - b > GAT_CTA_DEG) softmax_row<true>(p, r, b, e, threadIdx.x, GAT_THREADS, s_red, lane, warp);
    }
  }
}

// ---------------------------------------------------------------- out[j,h] = sum_k vals[eidx[k], h]   (d el)
struct GatSegSum {
  const int32_t* rowptr; const int32_t* eidx; const float* vals; float* out;
  int64_t n_rows; int32_t H;
};

template <bool CTA>
__device__ __forceinline__ void segsum_row(const GatSegSum& p, int64_t j, int b, int e, int tid, int nt,
                                           float (*s_red)[GAT_MAXH], int lane, int warp) {
  float s[GAT_MAXH];
#pragma unroll
  for (int h = 0; h < GAT_MAXH; ++h) s[h] = 0.f;
  for (int k = b + tid; k < e; k += nt) {
    const float* x = p.vals + (size_t)(p.eidx ? __ldg(p.eidx + k) : k) * p.H;
#pragma unroll
    for (int h = 0; h < GAT_MAXH; ++h)
      if (h < p.H) s[h] += __ldg(x + h);
  }
  reduce_heads<CTA, false>(s, p.H, s_red, lane, warp);
  if (tid == 0) {
#pragma unroll
    for (int h = 0; h < GAT_MAXH; ++h)
      if (h < p.H) p.out[(size_t)j * p.H + h] = s[h];
  }
}

__global__ void __launch_bounds__(GAT_THREADS) gat_segment_sum_kernel(const GatSegSum p) {
  __shared__ float s_red[GAT_WARPS][GAT_MAXH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t r0 = (int64_t)blockIdx.x * GAT_WARPS; r0 < p.n_rows; r0 += (int64_t)gridDim.x * GAT_WARPS) {
    const int64_t j = r0 + warp;
    if (j < p.n_rows) {
      const int b = __ldg(p.rowptr + j), e = __ldg(p.rowptr + j + 1);
      if (e - b <= GAT_CTA_DEG) segsum_row<false>(p, j, b, e, lane, 32, s_red, lane, warp);
    }
    for (int w = 0; w < GAT_WARPS; ++w) {
      const int64_t r = r0 + w;
      if (r >= p.n_rows) break;
      const int b = __ldg(p.rowptr + r), e = __ldg(p.rowptr + r + 1);
      if (e - b > GAT_CTA_DEG) segsum_row<true>(p, r, b, e, threadIdx.x, GAT_THREADS, s_red, lane, warp);
    }
  }
}

// ---------------------------------------------------------------- aggregate
struct GatAgg {
  const int32_t* rowptr; const int32_t* col; const int32_t* eidx;   // eidx: position of edge k in a[] (NULL: k)
  const int32_t* chunk_rowptr; const int32_t* hub_rows; const int32_t* hub_segptr;
  const float* a; const float* ft; float* out; float* ws;   // ws: [n_seg, K] hub-segment partials
  int64_t ldf, ldo;
  int32_t n_chunks, n_hub, n_seg, seg_len, hub_threshold, H, D, K;
};

// segment s of the hub plan -> (hub index, edge range)
__device__ __forceinline__ void hub_segment(const int32_t* rowptr, const int32_t* hub_rows, const int32_t* hub_segptr,
                                            int n_hub, int seg_len, int s, int& r, int& b, int& e) {
  int lo = 0, hi = n_hub;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(hub_segptr + mid) <= s) lo = mid; else hi = mid;
  }
  r = __ldg(hub_rows + lo);
  b = __ldg(rowptr + r) + (s - __ldg(hub_segptr + lo)) * seg_len;
  e = min(__ldg(rowptr + r + 1), b + seg_len);
}

// lane handles vectors v = lane + 32*j (j < NJ) of width W; head of vector v = (v*W)/D.  The group's warps take
// groups of U consecutive edges: warp `first` of `stride` warps starts at beg + first*U.
template <typename V, int NJ, int U>
__device__ __forceinline__ void agg_edges(const GatAgg& p, int beg, int end, int first, int stride, int lane, int nvec,
                                          const int (&head)[NJ], V (&acc)[NJ]) {
  constexpr int W = VecTraits<V>::W;
  const V* F = reinterpret_cast<const V*>(p.ft);
  const size_t ldv = (size_t)(p.ldf / W);
  for (int k0 = beg + first * U; k0 < end; k0 += stride * U) {
    V x[U][NJ];
    float w[U][NJ];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int kk = k0 + u;
      if (kk < end) {
        const V* row = F + (size_t)__ldg(p.col + kk) * ldv + lane;
        const float* ak = p.a + (size_t)(p.eidx ? __ldg(p.eidx + kk) : kk) * p.H;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          if (lane + 32 * j < nvec) { x[u][j] = vldg(row + 32 * j); w[u][j] = __ldg(ak + head[j]); }
          else { vzero(x[u][j]); w[u][j] = 0.f; }
        }
      } else {
#pragma unroll
        for (int j = 0; j < NJ; ++j) { vzero(x[u][j]); w[u][j] = 0.f; }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < NJ; ++j) vfma(acc[j], w[u][j], x[u][j]);
  }
}

template <typename V, int NJ, int U>
__global__ void __launch_bounds__(GAT_THREADS) gat_aggregate_kernel(const GatAgg p) {
  constexpr int W = VecTraits<V>::W;
  extern __shared__ float s_row[];   // K floats (hub-segment CTAs)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = p.K / W;
  int head[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) head[j] = ((lane + 32 * j) * W) / p.D;
  if ((int)blockIdx.x < p.n_seg) {     // hub segment: warps stride over groups of U edges, combined in warp order
    int r, b, e;
    hub_segment(p.rowptr, p.hub_rows, p.hub_segptr, p.n_hub, p.seg_len, blockIdx.x, r, b, e);
    V acc[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) vzero(acc[j]);
    agg_edges<V, NJ, U>(p, b, e, warp, GAT_WARPS, lane, nvec, head, acc);
    for (int i = threadIdx.x; i < p.K; i += GAT_THREADS) s_row[i] = 0.f;
    __syncthreads();
    V* sv = reinterpret_cast<V*>(s_row);
    for (int w = 0; w < GAT_WARPS; ++w) {
      if (warp == w) {
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          if (lane + 32 * j < nvec) { V t = sv[lane + 32 * j]; vadd(t, acc[j]); sv[lane + 32 * j] = t; }
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < p.K; i += GAT_THREADS) p.ws[(size_t)blockIdx.x * p.K + i] = s_row[i];
    return;
  }
  V* O = reinterpret_cast<V*>(p.out);
  const size_t ldov = (size_t)(p.ldo / W);
  const int chunk = (blockIdx.x - p.n_seg) * GAT_WARPS + warp;
  if (chunk >= p.n_chunks) return;
  const int r0 = __ldg(p.chunk_rowptr + chunk), r1 = __ldg(p.chunk_rowptr + chunk + 1);
  for (int r = r0; r < r1; ++r) {
    const int b = __ldg(p.rowptr + r), e = __ldg(p.rowptr + r + 1);
    if (e - b > p.hub_threshold) continue;
    V acc[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) vzero(acc[j]);
    agg_edges<V, NJ, U>(p, b, e, 0, 1, lane, nvec, head, acc);
#pragma unroll
    for (int j = 0; j < NJ; ++j)
      if (lane + 32 * j < nvec) O[(size_t)r * ldov + lane + 32 * j] = acc[j];
  }
}

// out[hub row] = its segments' partials added in segment order
__global__ void __launch_bounds__(GAT_THREADS) gat_hub_finalize_kernel(const int32_t* __restrict__ hub_rows,
                                                                       const int32_t* __restrict__ hub_segptr,
                                                                       const float* __restrict__ ws, float* __restrict__ out,
                                                                       int64_t ldo, int K) {
  const int r = __ldg(hub_rows + blockIdx.x);
  const int q0 = __ldg(hub_segptr + blockIdx.x), q1 = __ldg(hub_segptr + blockIdx.x + 1);
  for (int i = threadIdx.x; i < K; i += GAT_THREADS) {
    float t = 0.f;
    for (int q = q0; q < q1; ++q) t += ws[(size_t)q * K + i];
    out[(size_t)r * ldo + i] = t;
  }
}

// ---------------------------------------------------------------- backward, per destination row
struct GatBwd {
  const int32_t* rowptr; const int32_t* col; const float* a; const float* ft; const float* dout;
  const float* el; const float* er;
  const float* scale;   // [nnz,H] or NULL: attention dropout, the aggregation used a*scale (scale = keep/(1-p)); d a = d(a*scale)*scale
  float* dpre;   // [nnz,H]  out: d loss / d (el[src]+er[dst])
  float* der;    // [n_rows,H] out (may be NULL when there is no er)
  const int32_t* chunk_rowptr; const int32_t* hub_rows; const int32_t* hub_segptr;
  float* ws;     // [n_seg, H] hub-segment partials of S
  int64_t ldf, ldd, n_rows;
  int32_t n_chunks, n_hub, n_seg, seg_len, hub_threshold;
  int32_t H, D, K;
  float slope;
};

// phase 1 over a group's share of the row's edges: d a[k,h] = <ft[src,h,:], dout[i,h,:]> staged in dpre, and the
// group-local S[h] += a[k,h] * d a[k,h] (same value on every lane).  Per-head sums over the lanes: SEG = false
// reduces one head at a time over the whole warp (H * 5 shuffles per edge; right for few wide heads, e.g. the
// teacher's 3 x 250); SEG = true needs D/W a power of two <= 32, so a head is an aligned group of lph lanes of one
// vector index j, and a segmented xor-reduction costs NJ * log2(lph) shuffles (many narrow heads).
template <typename V, int NJ, int U, bool SEG>
__device__ __forceinline__ void bwd_edges(const GatBwd& p, int beg, int end, int first, int stride, int lane, int nvec,
                                          const V (&g)[NJ], const int (&head)[NJ], float (&S)[GAT_MAXH]) {
  constexpr int W = VecTraits<V>::W;
  const V* F = reinterpret_cast<const V*>(p.ft);
  const size_t ldv = (size_t)(p.ldf / W);
  const int lph = p.D / W;
  float Sj[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) Sj[j] = 0.f;
  for (int k0 = beg + first * U; k0 < end; k0 += stride * U) {
    V x[U][NJ];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int kk = k0 + u;
      if (kk < end) {
        const V* row = F + (size_t)__ldg(p.col + kk) * ldv + lane;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          if (lane + 32 * j < nvec) x[u][j] = vldg(row + 32 * j);
          else vzero(x[u][j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < NJ; ++j) vzero(x[u][j]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int kk = k0 + u;
      if (kk < end) {
        float d[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) d[j] = vdot(x[u][j], g[j]);
        if (SEG) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
              if (o < lph) d[j] += __shfl_xor_sync(FULL_MASK, d[j], o);
            if (lane + 32 * j < nvec) {
              const size_t o = (size_t)kk * p.H + head[j];
              const float da = p.scale ? d[j] * __ldg(p.scale + o) : d[j];
              if ((lane & (lph - 1)) == 0) p.dpre[o] = da;
              Sj[j] = fmaf(__ldg(p.a + o), da, Sj[j]);
            }
          }
        } else {
#pragma unroll
          for (int h = 0; h < GAT_MAXH; ++h)
            if (h < p.H) {
              float part = 0.f;
#pragma unroll
              for (int j = 0; j < NJ; ++j) part += (head[j] == h) ? d[j] : 0.f;
              const size_t o = (size_t)kk * p.H + h;
              const float da = p.scale ? gsum(part) * __ldg(p.scale + o) : gsum(part);
              if (lane == 0) p.dpre[o] = da;
              S[h] = fmaf(__ldg(p.a + o), da, S[h]);
            }
        }
      }
    }
  }
  if (SEG) {   // head h lives in vector index v = h*lph: lane v%32 of j = v/32 holds its S
#pragma unroll
    for (int h = 0; h < GAT_MAXH; ++h)
      if (h < p.H) {
        const int v = h * lph;
        float val = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) val = (j == (v >> 5)) ? Sj[j] : val;
        S[h] = __shfl_sync(FULL_MASK, val, v & 31);
      }
  }
}

// phase 2: d e = a (d a - S);  d pre = d e * leaky'(pre);  dr[h] = this thread's share of d er[i,h]
__device__ __forceinline__ void bwd_phase2(const GatBwd& p, int64_t i, int b, int e, int tid, int nt,
                                           const float (&S)[GAT_MAXH], float (&dr)[GAT_MAXH]) {
  float r[GAT_MAXH];
#pragma unroll
  for (int h = 0; h < GAT_MAXH; ++h) {
    r[h] = (p.er && h < p.H) ? __ldg(p.er + (size_t)i * p.H + h) : 0.f;
    dr[h] = 0.f;
  }
  for (int k = b + tid; k < e; k += nt) {
    const float* x = p.el + (size_t)__ldg(p.col + k) * p.H;
#pragma unroll
    for (int h = 0; h < GAT_MAXH; ++h)
      if (h < p.H) {
        const size_t o = (size_t)k * p.H + h;
        const float de = p.a[o] * (p.dpre[o] - S[h]);
        const float dp = (__ldg(x + h) + r[h]) > 0.f ? de : de * p.slope;
        p.dpre[o] = dp;
        dr[h] += dp;
      }
  }
}

template <typename V, int NJ, int U, bool SEG>
__global__ void __launch_bounds__(GAT_THREADS) gat_bwd_rows_kernel(const GatBwd p) {
  constexpr int W = VecTraits<V>::W;
  __shared__ float s_S[GAT_WARPS][GAT_MAXH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = p.K / W;
  const V* G = reinterpret_cast<const V*>(p.dout);
  const size_t lddv = (size_t)(p.ldd / W);
  int head[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) head[j] = ((lane + 32 * j) * W) / p.D;
  V g[NJ];
  float S[GAT_MAXH], dr[GAT_MAXH];
#pragma unroll
  for (int h = 0; h < GAT_MAXH; ++h) S[h] = 0.f;
  if ((int)blockIdx.x < p.n_seg) {     // hub segment: phase 1 only, partial S to the workspace
    int i, b, e;
    hub_segment(p.rowptr, p.hub_rows, p.hub_segptr, p.n_hub, p.seg_len, blockIdx.x, i, b, e);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      if (lane + 32 * j < nvec) g[j] = vldg(G + (size_t)i * lddv + lane + 32 * j);
      else vzero(g[j]);
    }
    bwd_edges<V, NJ, U, SEG>(p, b, e, warp, GAT_WARPS, lane, nvec, g, head, S);
    if (lane == 0) {
#pragma unroll
      for (int h = 0; h < GAT_MAXH; ++h)
        if (h < p.H) s_S[warp][h] = S[h];
    }
    __syncthreads();
    if (threadIdx.x < p.H) {
      float t = 0.f;
      for (int w = 0; w < GAT_WARPS; ++w) t += s_S[w][threadIdx.x];
      p.ws[(size_t)blockIdx.x * p.H + threadIdx.x] = t;
    }
    return;
  }
  const int chunk = (blockIdx.x - p.n_seg) * GAT_WARPS + warp;
  if (chunk >= p.n_chunks) return;
  const int r0 = __ldg(p.chunk_rowptr + chunk), r1 = __ldg(p.chunk_rowptr + chunk + 1);
  for (int i = r0; i < r1; ++i) {
    const int b = __ldg(p.rowptr + i), e = __ldg(p.rowptr + i + 1);
    if (e - b > p.hub_threshold) continue;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      if (lane + 32 * j < nvec) g[j] = vldg(G + (size_t)i * lddv + lane + 32 * j);
      else vzero(g[j]);
    }
#pragma unroll
    for (int h = 0; h < GAT_MAXH; ++h) S[h] = 0.f;
    bwd_edges<V, NJ, U, SEG>(p, b, e, 0, 1, lane, nvec, g, head, S);
    __syncwarp();                                   // lane 0's staged d a is read by every lane below
    bwd_phase2(p, i, b, e, lane, 32, S, dr);
    if (p.der) {
#pragma unroll
      for (int h = 0; h < GAT_MAXH; ++h)
        if (h < p.H) { const float t = gsum(dr[h]); if (lane == 0) p.der[(size_t)i * p.H + h] = t; }
    }
  }
}

// hub rows, after their segments: S = sum of the segment partials (segment order), then phase 2 over the whole row
__global__ void __launch_bounds__(GAT_THREADS) gat_bwd_hub_finalize_kernel(const GatBwd p) {
  __shared__ float s_S[GAT_WARPS][GAT_MAXH];
  __shared__ float s_tot[GAT_MAXH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = __ldg(p.hub_rows + blockIdx.x);
  const int b = __ldg(p.rowptr + i), e = __ldg(p.rowptr + i + 1);
  if (threadIdx.x < p.H) {
    float t = 0.f;
    for (int q = __ldg(p.hub_segptr + blockIdx.x); q < __ldg(p.hub_segptr + blockIdx.x + 1); ++q)
      t += p.ws[(size_t)q * p.H + threadIdx.x];
    s_tot[threadIdx.x] = t;
  }
  __syncthreads();
  float S[GAT_MAXH], dr[GAT_MAXH];
#pragma unroll
  for (int h = 0; h < GAT_MAXH; ++h) S[h] = h < p.H ? s_tot[h] : 0.f;
  bwd_phase2(p, i, b, e, threadIdx.x, GAT_THREADS, S, dr);
  if (p.der) {
#pragma unroll
    for (int h = 0; h < GAT_MAXH; ++h)
      if (h < p.H) { const float t = gsum(dr[h]); if (lane == 0) s_S[warp][h] = t; }
    __syncthreads();
    if (threadIdx.x < p.H) {
      float t = 0.f;
      for (int w = 0; w < GAT_WARPS; ++w) t += s_S[w][threadIdx.x];
      p.der[(size_t)i * p.H + threadIdx.x] = t;
    }
  }
}

static inline int rows_grid(int64_t n) {
  int64_t g = (n + GAT_WARPS - 1) / GAT_WARPS;
  if (g > 148 * 16) g = 148 * 16;
  return (int)(g < 1 ? 1 : g);
}

template <typename V, int NJ, int U>
static int launch_agg_nj(const GatAgg& p, cudaStream_t st) {
  int rc;
  const int grid = p.n_seg + (p.n_chunks + GAT_WARPS - 1) / GAT_WARPS;
  gat_aggregate_kernel<V, NJ, U><<<grid, GAT_THREADS, p.n_seg > 0 ? p.K * sizeof(float) : 0, st>>>(p);
  if ((rc = check_launch())) return rc;
  if (p.n_hub > 0) {
    gat_hub_finalize_kernel<<<p.n_hub, GAT_THREADS, 0, st>>>(p.hub_rows, p.hub_segptr, p.ws, p.out, p.ldo, p.K);
    if ((rc = check_launch())) return rc;
  }
  return B200GNN_OK;
}

template <typename V, int NJ, int U, bool SEG>
static int launch_bwd_seg(const GatBwd& p, cudaStream_t st) {
  int rc;
  const int grid = p.n_seg + (p.n_chunks + GAT_WARPS - 1) / GAT_WARPS;
  gat_bwd_rows_kernel<V, NJ, U, SEG><<<grid, GAT_THREADS, 0, st>>>(p);
  if ((rc = check_launch())) return rc;
  if (p.n_hub > 0) {
    gat_bwd_hub_finalize_kernel<<<p.n_hub, GAT_THREADS, 0, st>>>(p);
    if ((rc = check_launch())) return rc;
  }
  return B200GNN_OK;
}
template <typename V, int NJ, int U>
static int launch_bwd_nj(const GatBwd& p, cudaStream_t st) {
  constexpr int W = VecTraits<V>::W;
  if constexpr (W == 4) {   // segmented reduction when it is cheaper than H whole-warp sums
    const int lph = p.D / W, nj = (p.K / W + 31) / 32;
    int lg = 0;
    while ((1 << lg) < lph) ++lg;
    if (lph <= 32 && (lph & (lph - 1)) == 0 && nj * lg < p.H * 5) return launch_bwd_seg<V, NJ, U, true>(p, st);
  }
  return launch_bwd_seg<V, NJ, U, false>(p, st);
}

// vectors per lane -> (NJ, U): keep about 8 row vectors in flight per lane; the narrow vector types only get the
// two widest shapes (they are the odd-width fallbacks)
template <typename V>
static int launch_agg(const GatAgg& p, cudaStream_t st) {
  constexpr int W = VecTraits<V>::W;
  const int nj = (p.K / W + 31) / 32;
  if constexpr (W == 4) {
    if (nj <= 1) return launch_agg_nj<V, 1, 8>(p, st);
    if (nj <= 2) return launch_agg_nj<V, 2, 4>(p, st);
  }
  if (nj <= 4) return launch_agg_nj<V, 4, 2>(p, st);
  if constexpr (W == 4) {
    if (nj <= 8) return launch_agg_nj<V, 8, 1>(p, st);
  }
  return launch_agg_nj<V, GAT_MAXJ, 1>(p, st);
}
template <typename V>
static int launch_bwd(const GatBwd& p, cudaStream_t st) {
  constexpr int W = VecTraits<V>::W;
  const int nj = (p.K / W + 31) / 32;
  if constexpr (W == 4) {
    if (nj <= 1) return launch_bwd_nj<V, 1, 8>(p, st);
    if (nj <= 2) return launch_bwd_nj<V, 2, 4>(p, st);
  }
  if (nj <= 4) return launch_bwd_nj<V, 4, 2>(p, st);
  if constexpr (W == 4) {
    if (nj <= 8) return launch_bwd_nj<V, 8, 1>(p, st);
  }
  return launch_bwd_nj<V, GAT_MAXJ, 1>(p, st);
}

}  // namespace b200gnn

using namespace b200gnn;

extern "C" int b200gnn_gat_edge_softmax_f32(const int32_t* rowptr, const int32_t* col, const float* el, const float* er,
                                            int64_t n_rows, int64_t H, float negative_slope, float softmax_eps, float* a,
                                            const uint8_t* edge_keep, void* stream) {
  if (!rowptr || !el || !a || n_rows < 0 || H <= 0 || H > GAT_MAXH) return B200GNN_ERR_BAD_ARG;
  if (n_rows == 0) return B200GNN_OK;
  GatSoftmax p;
  p.rowptr = rowptr; p.col = col; p.el = el; p.er = er; p.a = a; p.keep = edge_keep; p.n_rows = n_rows; p.H = (int32_t)H;
  p.slope = negative_slope; p.eps = softmax_eps;
  gat_edge_softmax_kernel<<<rows_grid(n_rows), GAT_THREADS, 0, (cudaStream_t)stream>>>(p);
  return check_launch();
}

extern "C" int b200gnn_gat_aggregate_f32(const int32_t* rowptr, const int32_t* col, const int32_t* eidx, const float* a,
                                         const float* ft, int64_t ldf, float* out, int64_t ldo, int64_t n_rows, int64_t H,
                                         int64_t D, const int32_t* chunk_rowptr, int64_t n_chunks, int32_t hub_threshold,
                                         int32_t seg_len, const int32_t* hub_rows, const int32_t* hub_segptr, int64_t n_hub,
                                         int64_t n_seg, float* hub_workspace, void* stream) {
  const int64_t K = H * D;
  if (!rowptr || !a || !ft || !out || n_rows < 0 || H <= 0 || D <= 0 || H > GAT_MAXH || ldf < K || ldo < K || !chunk_rowptr ||
      n_chunks < 0 || n_hub < 0 || n_seg < n_hub ||
      (n_hub > 0 && (!hub_rows || !hub_segptr || !hub_workspace || seg_len <= 0)))
    return B200GNN_ERR_BAD_ARG;
  if (n_rows == 0 || n_chunks == 0) return B200GNN_OK;
  GatAgg p;
  p.rowptr = rowptr; p.col = col; p.eidx = eidx; p.chunk_rowptr = chunk_rowptr; p.hub_rows = hub_rows;
  p.hub_segptr = hub_segptr; p.ws = hub_workspace;
  p.a = a; p.ft = ft; p.out = out; p.ldf = ldf; p.ldo = ldo;
  p.n_chunks = (int32_t)n_chunks; p.n_hub = (int32_t)n_hub; p.n_seg = (int32_t)(n_hub > 0 ? n_seg : 0); p.seg_len = seg_len;
  p.hub_threshold = hub_threshold;
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
                                        int32_t seg_len, const int32_t* hub_rows, const int32_t* hub_segptr, int64_t n_hub,
                                        int64_t n_seg, float* hub_workspace, const float* attn_scale, void* stream) {
  const int64_t K = H * D;
  if (!rowptr || !a || !ft || !dout || !el || !dpre || n_rows < 0 || H <= 0 || D <= 0 || H > GAT_MAXH || ldf < K || ldd < K ||
      !chunk_rowptr || n_chunks < 0 || n_hub < 0 || n_seg < n_hub ||
      (n_hub > 0 && (!hub_rows || !hub_segptr || !hub_workspace || seg_len <= 0)))
    return B200GNN_ERR_BAD_ARG;
  if (n_rows == 0 || n_chunks == 0) return B200GNN_OK;
  GatBwd p;
  p.rowptr = rowptr; p.col = col; p.a = a; p.ft = ft; p.dout = dout; p.el = el; p.er = er; p.dpre = dpre; p.der = der;
  p.scale = attn_scale;
  p.chunk_rowptr = chunk_rowptr; p.hub_rows = hub_rows; p.hub_segptr = hub_segptr; p.ws = hub_workspace;
  p.ldf = ldf; p.ldd = ldd; p.n_rows = n_rows; p.n_chunks = (int32_t)n_chunks; p.n_hub = (int32_t)n_hub;
  p.n_seg = (int32_t)(n_hub > 0 ? n_seg : 0); p.seg_len = seg_len;
  p.hub_threshold = hub_threshold; p.H = (int32_t)H; p.D = (int32_t)D; p.K = (int32_t)K; p.slope = negative_slope;
  cudaStream_t st = (cudaStream_t)stream;
  if (D % 4 == 0 && ldf % 4 == 0 && ldd % 4 == 0 && aligned_to(ft, 16) && aligned_to(dout, 16) && K <= 4 * 32 * GAT_MAXJ)
    return launch_bwd<float4>(p, st);
  if (D % 2 == 0 && ldf % 2 == 0 && ldd % 2 == 0 && aligned_to(ft, 8) && aligned_to(dout, 8) && K <= 2 * 32 * GAT_MAXJ)
    return launch_bwd<float2>(p, st);
  if (K <= 32 * GAT_MAXJ) return launch_bwd<float>(p, st);
  return B200GNN_ERR_UNSUPPORTED;
}

extern "C" int b200gnn_segment_sum_heads_f32(const int32_t* rowptr, const int32_t* eidx, const float* vals, int64_t n_rows,
                                             int64_t H, float* out, void* stream) {
  if (!rowptr || !vals || !out || n_rows < 0 || H <= 0 || H > GAT_MAXH) return B200GNN_ERR_BAD_ARG;
  if (n_rows == 0) return B200GNN_OK;
  GatSegSum p;
  p.rowptr = rowptr; p.eidx = eidx; p.vals = vals; p.out = out; p.n_rows = n_rows; p.H = (int32_t)H;
  gat_segment_sum_kernel<<<rows_grid(n_rows), GAT_THREADS, 0, (cudaStream_t)stream>>>(p);
  return check_launch();
}
