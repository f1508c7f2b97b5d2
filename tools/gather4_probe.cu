// Micro-probe: TMA tile::gather4 (UTMALDG.2D.GATHER4) as the SpMM neighbour-row fetch — one instruction brings FOUR rows
// (given by four row indices) of a [n_rows, 256] fp32 matrix, W columns each, into contiguous shared memory.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/gather4_probe.cu -o /tmp/gather4_probe && /tmp/gather4_probe
// 1. semantics: which boxDim encoding the instruction wants ({W,1} or {W,4}) and the smem layout it produces;
// 2. throughput of random 4-row gathers for slab widths W = 64 / 128 / 256 floats over the ARXIV-shape operand.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
  return reinterpret_cast<EncodeTiledFn>(p);
}
static bool make_map(CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_w, uint32_t box_h) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {box_w, box_h};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void gather4(void* dst, const CUtensorMap* tm, int col, int r0, int r1, int r2, int r3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
      : "memory");
}

// ---- 1. semantics
__global__ void sem_kernel(const __grid_constant__ CUtensorMap tm, int W, int col0, int4 rows, uint32_t* out) {
  extern __shared__ __align__(128) uint32_t buf[];
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < 4 * W + 64; i += blockDim.x) buf[i] = 0xdeadbeefu;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, 4 * W * 4);
    gather4(buf, &tm, col0, rows.x, rows.y, rows.z, rows.w, &bar);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 4 * W + 64; i += blockDim.x) out[i] = buf[i];
}

// ---- 2. throughput: per-warp ring of D slots (each 4 rows x W floats), one mbarrier per slot group of G gathers
template <int W, int RING, int G>
__global__ void __launch_bounds__(256) g4_gather(const __grid_constant__ CUtensorMap tm, const int* __restrict__ idx, int n_idx,
                                                 int col0, float4* __restrict__ sink) {
  constexpr int SLOT_B = 4 * W * 4, D = RING / SLOT_B, NG = D / G;
  constexpr int VPR = W / 4;                       // float4 per row
  extern __shared__ __align__(128) char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  char* ring = smem + (size_t)warp * RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)(blockDim.x >> 5) * RING) + warp * NG;
  if (lane < NG) mbar_init(bars + lane, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int per = (n_idx + nw - 1) / nw / 32 * 32;
  const int lo = min(n_idx, gw * per), hi = min(n_idx, lo + per);
  const int n = hi - lo;                           // multiple of 32 (edges); one gather4 = 4 edges
  const int nq = n / 4;                            // gathers
  float4 acc = make_float4(0, 0, 0, 0);
  uint32_t phases = 0;
  int cI = 0;
  auto issue = [&](int q0, int slot0, int b) {     // gathers q0..q0+G-1
    if (q0 >= nq) return;
    if (((q0 * 4) & 31) == 0) cI = __ldg(idx + lo + q0 * 4 + lane);
    const int cnt = min(G, nq - q0);
    if (lane == 0) mbar_expect_tx(bars + b, cnt * SLOT_B);
    // lane 4m (m = gather index inside the 32-edge window) collects the four row ids
    const int r1 = __shfl_down_sync(0xffffffffu, cI, 1), r2 = __shfl_down_sync(0xffffffffu, cI, 2),
              r3 = __shfl_down_sync(0xffffffffu, cI, 3);
    const int u = (lane >> 2) - (q0 & 7);
    if ((lane & 3) == 0 && u >= 0 && u < cnt) gather4(ring + (slot0 + u) * SLOT_B, &tm, col0, cI, r1, r2, r3, bars + b);
  };
#pragma unroll
  for (int g = 0; g < NG; ++g) issue(g * G, g * G, g);
  int b = 0;
  for (int q = 0, s0 = 0; q < nq; q += G) {
    mbar_wait(bars + b, (phases >> b) & 1);
    phases ^= 1u << b;
    const int cnt = min(G, nq - q);
    for (int u = 0; u < cnt; ++u) {
      const float4* s = reinterpret_cast<const float4*>(ring + (s0 + u) * SLOT_B);
#pragma unroll
      for (int v = lane; v < 4 * VPR; v += 32) {
        const float4 x = s[v];
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
      }
    }
    __syncwarp();
    issue(q + D, s0, b);
    s0 = (s0 + G) & (D - 1);
    b = b + 1 == NG ? 0 : b + 1;
  }
  if (acc.x == 123.456f) sink[0] = acc;
}

static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }

template <int W, int RING, int G>
static void run(const CUtensorMap& tm, const int* d_idx, int n_idx, float4* sink, int ctas_per_sm, int box_h) {
  const int smem = 8 * RING + 8 * (RING / (16 * W) / G) * 8 + 64;
  cudaFuncSetAttribute(g4_gather<W, RING, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  g4_gather<W, RING, G><<<148 * ctas_per_sm, 256, smem>>>(tm, d_idx, n_idx, 0, sink);
  cudaEventRecord(a);
  for (int it = 0; it < 5; ++it) g4_gather<W, RING, G><<<148 * ctas_per_sm, 256, smem>>>(tm, d_idx, n_idx, 0, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
  cudaError_t e = cudaGetLastError();
  printf("{\"probe\":\"gather4\",\"slab_floats\":%d,\"footprint_MB\":%.0f,\"ring_bytes\":%d,\"group\":%d,\"ctas_per_sm\":%d,"
         "\"box_h\":%d,\"GBps\":%.1f,\"Grows_per_s\":%.2f,\"err\":\"%s\"}\n", W, 169343.0 * W * 4 / 1048576, RING, G, ctas_per_sm,
         box_h, (double)n_idx * W * 4 / ms / 1e6, (double)n_idx / ms / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
  fflush(stdout);
}

int main() {
  const int n_rows = 169343, K = 256;
  std::vector<uint32_t> h((size_t)n_rows * K);
  for (int r = 0; r < n_rows; ++r) for (int c = 0; c < K; ++c) h[(size_t)r * K + c] = ((uint32_t)r << 8) | (uint32_t)c;
  float* X; cudaMalloc(&X, h.size() * 4); cudaMemcpy(X, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  uint32_t* d_out; cudaMalloc(&d_out, (4 * 256 + 64) * 4);
  std::vector<uint32_t> out(4 * 256 + 64);

  int good_box_h = 0;
  for (int box_h : {1}) {   // {W,4} boxes are rejected by the instruction (illegal instruction): the box is ONE row
    for (int W : {64, 128, 256}) {
      CUtensorMap tm;
      if (!make_map(&tm, X, n_rows, K, K, W, box_h)) { printf("{\"probe\":\"gather4_sem\",\"box_h\":%d,\"W\":%d,\"encode\":\"failed\"}\n", box_h, W); continue; }
      const int4 rows = make_int4(5, 169000, 77, 12345);
      const int col0 = (W == 256) ? 0 : W;       // second slab
      cudaMemset(d_out, 0, out.size() * 4);
      sem_kernel<<<1, 128, (4 * W + 64) * 4>>>(tm, W, col0, rows, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("{\"probe\":\"gather4_sem\",\"box_h\":%d,\"W\":%d,\"error\":\"%s\"}\n", box_h, W, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
      const int rr[4] = {rows.x, rows.y, rows.z, rows.w};
      int bad = 0;
      for (int j = 0; j < 4; ++j) for (int c = 0; c < W; ++c) bad += out[j * W + c] != (((uint32_t)rr[j] << 8) | (uint32_t)(col0 + c));
      int tail_ok = 1;
      for (int i = 4 * W; i < 4 * W + 64; ++i) tail_ok &= out[i] == 0xdeadbeefu;
      printf("{\"probe\":\"gather4_sem\",\"box_h\":%d,\"W\":%d,\"mismatches\":%d,\"tail_untouched\":%d,\"first\":[%u,%u,%u,%u]}\n", box_h, W, bad,
             tail_ok, out[0] >> 8, out[W] >> 8, out[2 * W] >> 8, out[3 * W] >> 8);
      if (!bad && !good_box_h) good_box_h = box_h;
    }
  }
  fflush(stdout);
  if (!good_box_h) return 0;

  const int n_idx = 4 << 20;
  std::vector<int> hi(n_idx);
  uint32_t s = 12345;
  for (int i = 0; i < n_idx; ++i) hi[i] = lcg(s) % n_rows;
  int* d; cudaMalloc(&d, n_idx * 4); cudaMemcpy(d, hi.data(), n_idx * 4, cudaMemcpyHostToDevice);
  float4* sink; cudaMalloc(&sink, 64);
  for (int cps : {2, 3}) {
    { CUtensorMap tm; make_map(&tm, X, n_rows, K, K, 64, good_box_h);
      run<64, 8192, 2>(tm, d, n_idx, sink, cps, good_box_h); run<64, 8192, 4>(tm, d, n_idx, sink, cps, good_box_h);
      run<64, 4096, 2>(tm, d, n_idx, sink, cps, good_box_h); run<64, 16384, 4>(tm, d, n_idx, sink, cps > 2 ? 2 : cps, good_box_h); }
    { CUtensorMap tm; make_map(&tm, X, n_rows, K, K, 128, good_box_h);
      run<128, 8192, 2>(tm, d, n_idx, sink, cps, good_box_h); run<128, 8192, 1>(tm, d, n_idx, sink, cps, good_box_h);
      run<128, 16384, 2>(tm, d, n_idx, sink, cps > 2 ? 2 : cps, good_box_h); run<128, 16384, 4>(tm, d, n_idx, sink, cps > 2 ? 2 : cps, good_box_h); }
    { CUtensorMap tm; make_map(&tm, X, n_rows, K, K, 256, good_box_h);
      run<256, 8192, 1>(tm, d, n_idx, sink, cps, good_box_h); run<256, 16384, 2>(tm, d, n_idx, sink, cps > 2 ? 2 : cps, good_box_h);
      run<256, 16384, 1>(tm, d, n_idx, sink, cps > 2 ? 2 : cps, good_box_h); }
  }
  return 0;
}
