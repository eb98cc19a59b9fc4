// Microbenchmark: issue cost (cycles per warp instruction, per SM sub-partition) of the conversions and special functions
// the softmax / activation epilogues are made of.   nvcc -arch=sm_100a -O3 -o xu_bench xu_bench.cu && ./xu_bench
// One CTA per SM, W warps per CTA (W/4 per sub-partition); every thread runs a dependent-free unrolled stream of N ops.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ void body(float (&v)[8], uint32_t (&u)[8]) {
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    if (MODE == 0) {          // ex2 only
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i + 1]));
    } else if (MODE == 1) {   // cvt f16x2 only
      asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(v[i]), "f"(v[i + 1]));
      asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u[i + 1]) : "f"(v[i + 1]), "f"(v[i]));
    } else if (MODE == 2) {   // the split as shipped: 2 ex2 + cvt hi + unpack + sub2 + cvt lo
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i + 1]));
      uint32_t hi, lo;
      asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v[i + 1]), "f"(v[i]));
      const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
      const float r0 = v[i] - h.x, r1 = v[i + 1] - h.y;
      asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
      u[i] ^= hi; u[i + 1] ^= lo;
    } else if (MODE == 3) {   // 2 ex2 + truncating split on the integer / FMA pipes (no cvt)
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i + 1]));
      const float s = 1.925929944387236e-34f;   // 2^-112
      const float S = 5.192296858534828e+33f;   // 2^112
      const uint32_t t0 = __float_as_uint(v[i] * s) & 0xffffe000u, t1 = __float_as_uint(v[i + 1] * s) & 0xffffe000u;
      const uint32_t hi = (t0 >> 13) | ((t1 >> 13) << 16);
      const float r0 = v[i] - __uint_as_float(t0) * S, r1 = v[i + 1] - __uint_as_float(t1) * S;
      const uint32_t lo = (__float_as_uint(r0 * s) >> 13) | ((__float_as_uint(r1 * s) >> 13) << 16);
      u[i] ^= hi; u[i + 1] ^= lo;
    } else if (MODE == 4) {   // unpack only: half2 -> float2
      const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&u[i]));
      v[i] += h.x; v[i + 1] += h.y;
    } else if (MODE == 5) {   // rcp only
      asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i + 1]));
    }
  }
}

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float v[8]; uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = 0.001f * (threadIdx.x + i); u[i] = threadIdx.x * 2654435761u + i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) body<MODE>(v, u);
  const long long t1 = clock64();
  float s = 0; uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += v[i]; x ^= u[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(x & 0x3fffffu);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_pair_note) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int warps : {4, 8, 16}) {
    const int iters = 2000;
    k<MODE><<<148, warps * 32, 0>>>(out, cyc, iters);
    k<MODE><<<148, warps * 32, 0>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    // 4 pairs per iteration per thread; warps/4 warps per sub-partition
    printf("%-44s warps/SMSP %d: %.1f cycles per PAIR of elements per sub-partition-warp\n", name, warps / 4, avg / iters / 4 / (warps / 4));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("2 x ex2.approx", 0);
  run<5>("2 x rcp.approx", 0);
  run<1>("2 x cvt.rn.satfinite.f16x2.f32", 0);
  run<4>("half2 -> float2 unpack + 2 fadd", 0);
  run<2>("2 ex2 + fp16 hi/lo split via cvt (shipped)", 0);
  run<3>("2 ex2 + truncating split on int/FMA pipes", 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
