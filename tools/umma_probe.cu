// Standalone probe of tcgen05.mma operand-descriptor semantics (sm_100a).  One CTA, one MMA (M=128, N=32, K=8, tf32),
// operands written to shared memory by ordinary threads in a software-swizzled layout, D read back and checked.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -I efficient-gnns_b200/csrc -I include tools/umma_probe.cu -o /tmp/umma_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace b200gnn::tc;

struct Variant {
  int mn_major;        // 0: K-major operands, 1: MN-major SW128 (16B atoms), 2: MN-major SW128 with 32B atoms (BASE32B)
  int group_stride;    // bytes between 32-element MN groups (MN-major) -- layout AND (by default) LBO
  int lbo, sbo;        // descriptor fields in bytes
  int a_major_bit, b_major_bit;
};

__global__ void probe(Variant v, float* Dout) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* As = smem;            // up to 16 KB
  uint8_t* Bs = smem + 16384;    // up to 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 32768 + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32768 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  __syncthreads();
  const int M = 128, N = 32, K = 8;
  // A[m][k] = (m+1) * 0.5 + k * 0.125 ; B[n][k] = (n+1) * 0.25 - k * 0.5   (exact in tf32)
  for (int i = tid; i < M * K; i += blockDim.x) {
    const int m = i / K, k = i % K;
    const float val = (m + 1) * 0.5f + k * 0.125f;
    uint32_t off;
    if (v.mn_major == 2) {
      const int g = m / 32, chunk = (m % 32) / 8, e = m % 8;
      off = g * v.group_stride + k * 128 + ((chunk ^ (k % 4)) * 32) + e * 4;
    } else if (v.mn_major) {
      const int g = m / 32, chunk = (m % 32) / 4, e = m % 4;
      off = g * v.group_stride + k * 128 + ((chunk ^ (k % 8)) * 16) + e * 4;
    } else {  // K-major SW128: row m (128 B), 8-row groups 1024 B apart; element k at byte k*4 -> chunk k/4
      const int chunk = k / 4, e = k % 4;
      off = (m / 8) * 1024 + (m % 8) * 128 + ((chunk ^ (m % 8)) * 16) + e * 4;
    }
    *reinterpret_cast<float*>(As + off) = val;
  }
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    const float val = (n + 1) * 0.25f - k * 0.5f;
    uint32_t off;
    if (v.mn_major == 2) {
      const int g = n / 32, chunk = (n % 32) / 8, e = n % 8;
      off = g * v.group_stride + k * 128 + ((chunk ^ (k % 4)) * 32) + e * 4;
    } else if (v.mn_major) {
      const int g = n / 32, chunk = (n % 32) / 4, e = n % 4;
      off = g * v.group_stride + k * 128 + ((chunk ^ (k % 8)) * 16) + e * 4;
    } else {
      const int chunk = k / 4, e = k % 4;
      off = (n / 8) * 1024 + (n % 8) * 128 + ((chunk ^ (n % 8)) * 16) + e * 4;
    }
    *reinterpret_cast<float*>(Bs + off) = val;
  }
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (tid == 32) {
    auto mk = [&](uint32_t saddr) {
      uint64_t d = 0;
      d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
      d |= (uint64_t)(v.lbo >> 4) << 16;
      d |= (uint64_t)(v.sbo >> 4) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)(v.mn_major == 2 ? 1 : 2) << 61;   // 1 = SWIZZLE_128B_BASE32B, 2 = SWIZZLE_128B
      return d;
    };
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)v.a_major_bit << 15) |
                           ((uint32_t)v.b_major_bit << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    mma_tf32(tmem, mk(smem_u32(As)), mk(smem_u32(Bs)), idesc, 0);
    mma_commit(bar);
  }
  if (warp >= 4 && warp < 8) {
    mbar_wait(bar, 0);
    tc_fence_after();
    const int q = warp & 3, lane = tid & 31;
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16), r);
    for (int j = 0; j < 32; ++j) Dout[(q * 32 + lane) * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64));
  }
}

int main() {
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  float* d; cudaMalloc(&d, 128 * 32 * 4);
  float* h = (float*)malloc(128 * 32 * 4);
  Variant vs[] = {
      {0, 0, 16, 1024, 0, 0},        // K-major reference (known good from gemm_tf32x3.cu)
      {1, 2048, 2048, 1024, 1, 1},   // MN-major as in gemm_wgrad: groups 2048 apart, LBO=2048, SBO=1024
      {1, 2048, 1024, 2048, 1, 1},   // LBO/SBO swapped
      {1, 1024, 1024, 1024, 1, 1},   // groups adjacent (CUTLASS atom order), LBO=1024
      {1, 1024, 1024, 4096, 1, 1},   // same, SBO=4096
      {1, 1024, 4096, 1024, 1, 1},   // groups adjacent, LBO/SBO swapped
      {1, 2048, 2048, 1024, 0, 0},   // MN layout but major bits clear (expect wrong)
      {2, 2048, 2048, 512, 1, 1},    // BASE32B: 16 rows per group, LBO = group stride, SBO = 4-row atom stride
      {2, 2048, 512, 2048, 1, 1},    // swapped
      {2, 1024, 1024, 512, 1, 1},    // 8 rows per group
      {2, 1024, 512, 1024, 1, 1},    // swapped
  };
  for (auto& v : vs) {
    cudaMemset(d, 0xff, 128 * 32 * 4);
    probe<<<1, 256, 40000>>>(v, d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 128 * 32 * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0; int nz = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 32; ++n) {
        double ref = 0;
        for (int k = 0; k < 8; ++k) ref += ((m + 1) * 0.5 + k * 0.125) * ((n + 1) * 0.25 - k * 0.5);
        maxerr = fmax(maxerr, fabs(h[m * 32 + n] - ref)); maxref = fmax(maxref, fabs(ref));
        nz += h[m * 32 + n] != 0.f;
      }
    printf("variant mn=%d gstride=%d lbo=%d sbo=%d majors=%d%d : err=%s maxerr=%.4g (ref max %.4g) nonzero=%d D[0][0..3]=%g %g %g %g D[1][0]=%g D[33][0]=%g\n",
           v.mn_major, v.group_stride, v.lbo, v.sbo, v.a_major_bit, v.b_major_bit, cudaGetErrorString(e), maxerr, maxref, nz,
           h[0], h[1], h[2], h[3], h[32], h[33 * 32]);
  }
  return 0;
}
