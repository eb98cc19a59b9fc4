// Microbenchmark (B200): how fast can every SM stream the SAME weight blob out of L2 into shared memory with
// cp.async.bulk, (a) unicast, one CTA per SM, (b) 2- or 4-CTA clusters where each CTA issues 1/csz of the tile loads and
// multicasts them to the whole cluster.  The decoder's chained kernels stream 2-3.5 MB of weight tiles per 128-point
// tile through a 3-4 slot ring; this measures the ceiling of that stream (no MMA: the consumer frees a slot at once).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_stream_bench l2_stream_bench.cu && ./l2_stream_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) if (++spins > (1u << 24)) { printf("timeout block %d bar %x\n", blockIdx.x, bar); __trap(); }
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// warp 0 lane 0 = producer, warp 1 lane 0 = consumer
template <int CSZ>
__global__ void stream_kernel(const uint8_t* blob, int tiles_per_pass, int passes, int slots, uint32_t tile_bytes, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t bars = base + slots * 32768;
  const uint32_t rank = CSZ > 1 ? cluster_rank() : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < slots; ++i) { mbar_init(bars + 8 * i, 1); mbar_init(bars + 64 + 8 * i, CSZ); }   // full: 1 arrive (+tx); empty: one per CTA
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (CSZ > 1) cluster_sync();
  const int total = tiles_per_pass * passes;
  if (threadIdx.x == 0) {
    int idx = 0; uint32_t ph = 0;
    for (int i = 0; i < total; ++i) {
      mbar_wait(bars + 64 + 8 * idx, ph ^ 1);                  // slot free in every CTA of the cluster
      mbar_expect_tx(bars + 8 * idx, tile_bytes);
      const uint8_t* src = blob + (size_t)(i % tiles_per_pass) * 32768;
      if (CSZ == 1) bulk_g2s(base + idx * 32768, src, tile_bytes, bars + 8 * idx);
      else if ((uint32_t)(i % CSZ) == rank) bulk_g2s_mc(base + idx * 32768, src, tile_bytes, bars + 8 * idx, (uint16_t)((1u << CSZ) - 1));
      if (++idx == slots) { idx = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int idx = 0; uint32_t ph = 0;
    unsigned long long acc = 0;
    for (int i = 0; i < total; ++i) {
      mbar_wait(bars + 8 * idx, ph);
      acc += *reinterpret_cast<volatile uint32_t*>(smem + (base - smem_u32(smem)) + idx * 32768 + (i & 63) * 4);
      if (CSZ == 1) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 64 + 8 * idx) : "memory");
      else for (uint32_t c = 0; c < (uint32_t)CSZ; ++c) mbar_arrive_cluster(bars + 64 + 8 * idx, c);
      if (++idx == slots) { idx = 0; ph ^= 1; }
    }
    if (acc == 0x1234567) sink[0] = acc;
  }
  __syncthreads();
  if (CSZ > 1) cluster_sync();
}

template <int CSZ>
static float run(const uint8_t* blob, int tiles, int passes, int slots, uint32_t tile_bytes, unsigned long long* sink, int grid) {
  size_t smem = (size_t)slots * 32768 + 1024 + 256;
  cudaFuncSetAttribute(stream_kernel<CSZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CSZ; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, stream_kernel<CSZ>, blob, tiles, passes, slots, tile_bytes, sink);
    cudaEventRecord(e1);
    if (err != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  return best;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int tiles = 64;                       // 2 MB blob (the MLP kernel's weight stream of one tile), L2 resident
  uint8_t* blob; cudaMalloc(&blob, (size_t)tiles * 32768); cudaMemset(blob, 1, (size_t)tiles * 32768);
  unsigned long long* sink; cudaMalloc(&sink, 8);
  const int passes = 40;
  printf("SMs %d; every CTA streams %d x 32 KB x %d passes\n", sms, tiles, passes);
  for (uint32_t tb : {32768u, 16384u}) {
    for (int slots : {3, 4, 6}) {
      const double bytes = (double)tiles * passes * tb;
      float a = run<1>(blob, tiles, passes, slots, tb, sink, sms);
      float b = run<2>(blob, tiles, passes, slots, tb, sink, sms / 2 * 2);
      float c = run<4>(blob, tiles, passes, slots, tb, sink, sms / 4 * 4);
      printf("tile %5u B, %d slots: unicast %.3f ms = %.1f GB/s/SM (%.2f TB/s chip) | cluster2 multicast %.3f ms = %.1f GB/s/SM delivered | cluster4 %.3f ms = %.1f GB/s/SM\n",
             tb, slots, a, bytes / a / 1e6, bytes * sms / a / 1e9, b, bytes / b / 1e6, c, bytes / c / 1e6);
    }
  }
  return 0;
}
