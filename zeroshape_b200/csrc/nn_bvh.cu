// Exact nearest neighbours through a flat bounding-box hierarchy -- the brute-force pose-search evaluator
// (reference utils/eval_3D.py:140-207: 6912 rotations x one bidirectional Chamfer of 10,000 x 10,000 points each =
// 1.4e12 pair evaluations per shape through external/chamfer3D/chamfer3D.cu:12-134).
//
// Same results as zs_chamfer_nn_fwd -- squared distance with the reference kernel's arithmetic (sqdist_ref below,
// bit-identical), lowest index on ties -- at ~1/20 of the pair evaluations:
//   build : one CTA per point set.  Points are sorted along a 30-bit Morton curve of the set's bounding cube
//           (cub::BlockRadixSort in shared memory) and cut into clusters of 32 consecutive points, each with its
//           tight axis-aligned box; 16 consecutive clusters share a super-box.  A surface cloud of 10,000 points gives
//           313 compact patches under 20 super-boxes.
//   query : one thread per query point, the target set's boxes in shared memory (every thread of the CTA scans the same
//           super-box list: broadcast reads, no divergence).  Pass 1 descends to the box with the smallest lower bound
//           and scans its 32 points; pass 2 opens every super-box, and inside it every box, whose (conservatively
//           rounded) lower bound does not exceed the best distance so far.  Exact: a skipped box cannot hold a closer --
//           or an equally close, lower-index -- point.
// Distances are evaluated on the caller's own fp32 coordinates (no re-centring / rotation inside), so they are the very
// numbers the brute-force kernel produces.
#include <cub/block/block_radix_sort.cuh>
#include "common.cuh"

namespace zs {

constexpr int BVH_THREADS = 1024, BVH_ITEMS = 16, BVH_MAX_N = BVH_THREADS * BVH_ITEMS;   // 16,384 points per set
constexpr int BVH_CLUSTER = 32, BVH_SUPER = 16;      // points per box, boxes per super-box (= one warp of the build kernel)

__device__ __forceinline__ float sqdist_ref_bvh(float tx, float ty, float tz, float qx, float qy, float qz) {
  // == sqdist_ref of chamfer.cu: what nvcc emits for the reference expression x2*x2+y2*y2+z2*z2 (chamfer3D.cu:32)
  float x2 = tx - qx, y2 = ty - qy, z2 = tz - qz;
  return __fmaf_rn(z2, z2, __fmaf_rn(x2, x2, __fmul_rn(y2, y2)));
}

__host__ __device__ inline int bvh_padded(int n) { return (n + BVH_CLUSTER - 1) / BVH_CLUSTER * BVH_CLUSTER; }
__host__ __device__ inline int bvh_supers(int n) { return (bvh_padded(n) / BVH_CLUSTER + BVH_SUPER - 1) / BVH_SUPER; }
// per set: float4 points[NP] (x, y, z, original index as int bits; padding = +inf), float boxes[NC][8] (min xyz, -, max xyz, -),
// float superboxes[NS][8]
__host__ __device__ inline size_t bvh_set_bytes(int n) {
  const size_t np = (size_t)bvh_padded(n);
  return np * 16 + (np / BVH_CLUSTER) * 32 + (size_t)bvh_supers(n) * 32;
}

__device__ __forceinline__ unsigned spread10(unsigned v) {     // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

using BvhSort = cub::BlockRadixSort<unsigned, BVH_THREADS, BVH_ITEMS, int>;

__global__ void __launch_bounds__(BVH_THREADS) nn_bvh_build_kernel(const float* __restrict__ pts, int n, uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  typename BvhSort::TempStorage& sort_tmp = *reinterpret_cast<typename BvhSort::TempStorage*>(sm_raw);
  __shared__ float red[6][32];
  __shared__ float bb[6];
  const int set = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const float* P = pts + (int64_t)set * n * 3;
  uint8_t* base = out + (size_t)set * bvh_set_bytes(n);
  const int NP = bvh_padded(n);
  float4* sorted = reinterpret_cast<float4*>(base);
  float* boxes = reinterpret_cast<float*>(base + (size_t)NP * 16);

  // bounding box of the set
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = t; i < n; i += BVH_THREADS) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { const float v = P[i * 3 + a]; mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v); }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
    if (lane == 0) { red[a][warp] = mn[a]; red[3 + a][warp] = mx[a]; }
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float v = red[a][lane], w = red[3 + a][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o)); w = fmaxf(w, __shfl_xor_sync(0xffffffffu, w, o)); }
      if (lane == 0) { bb[a] = v; bb[3 + a] = w; }
    }
  }
  __syncthreads();
  const float ext = fmaxf(fmaxf(bb[3] - bb[0], bb[4] - bb[1]), fmaxf(bb[5] - bb[2], 1e-30f));
  const float inv = 1024.0f / ext;

  // Morton keys (blocked arrangement: thread t owns items t*ITEMS .. +ITEMS), padding sorts last
  unsigned keys[BVH_ITEMS];
  int vals[BVH_ITEMS];
#pragma unroll
  for (int k = 0; k < BVH_ITEMS; ++k) {
    const int i = t * BVH_ITEMS + k;
    vals[k] = i;
    if (i < n) {
      const float x = (P[i * 3] - bb[0]) * inv, y = (P[i * 3 + 1] - bb[1]) * inv, z = (P[i * 3 + 2] - bb[2]) * inv;
      const unsigned ix = (unsigned)fminf(fmaxf(x, 0.f), 1023.f), iy = (unsigned)fminf(fmaxf(y, 0.f), 1023.f),
                     iz = (unsigned)fminf(fmaxf(z, 0.f), 1023.f);
      keys[k] = spread10(ix) | (spread10(iy) << 1) | (spread10(iz) << 2);
    } else {
      keys[k] = 0xffffffffu;
    }
  }
  BvhSort(sort_tmp).Sort(keys, vals, 0, 30 + 2);       // 30 key bits + the all-ones padding key
  // sorted position t*ITEMS + k holds original point vals[k]; one cluster = 32 positions = the items of threads 2c, 2c+1
  float cmn[3] = {INFINITY, INFINITY, INFINITY}, cmx[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int k = 0; k < BVH_ITEMS; ++k) {
    const int pos = t * BVH_ITEMS + k;
    if (pos >= NP) continue;
    const int i = vals[k];
    float4 v = make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(0x7fffffff));
    if (i < n && keys[k] != 0xffffffffu) {
      v = make_float4(P[i * 3], P[i * 3 + 1], P[i * 3 + 2], __int_as_float(i));
      cmn[0] = fminf(cmn[0], v.x); cmn[1] = fminf(cmn[1], v.y); cmn[2] = fminf(cmn[2], v.z);
      cmx[0] = fmaxf(cmx[0], v.x); cmx[1] = fmaxf(cmx[1], v.y); cmx[2] = fmaxf(cmx[2], v.z);
    }
    sorted[pos] = v;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    cmn[a] = fminf(cmn[a], __shfl_xor_sync(0xffffffffu, cmn[a], 1));
    cmx[a] = fmaxf(cmx[a], __shfl_xor_sync(0xffffffffu, cmx[a], 1));
  }
  const int c = t >> 1;
  if ((t & 1) == 0 && c < NP / BVH_CLUSTER) {
    float* b = boxes + (size_t)c * 8;
    b[0] = cmn[0]; b[1] = cmn[1]; b[2] = cmn[2]; b[3] = 0.f;
    b[4] = cmx[0]; b[5] = cmx[1]; b[6] = cmx[2]; b[7] = 0.f;
  }
  // a warp holds 16 consecutive clusters = one super-box (empty clusters contribute +inf / -inf)
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 1; o >>= 1) {
      cmn[a] = fminf(cmn[a], __shfl_xor_sync(0xffffffffu, cmn[a], o));
      cmx[a] = fmaxf(cmx[a], __shfl_xor_sync(0xffffffffu, cmx[a], o));
    }
  }
  if (lane == 0 && warp < bvh_supers(n)) {
    float* b = boxes + (size_t)(NP / BVH_CLUSTER) * 8 + (size_t)warp * 8;
    b[0] = cmn[0]; b[1] = cmn[1]; b[2] = cmn[2]; b[3] = 0.f;
    b[4] = cmx[0]; b[5] = cmx[1]; b[6] = cmx[2]; b[7] = 0.f;
  }
}

constexpr int BVQ_THREADS = 256;

__global__ void __launch_bounds__(BVQ_THREADS) nn_bvh_query_kernel(const uint8_t* __restrict__ bvh, int sets_t, int n,
                                                                   const float* __restrict__ q, int sets_q, int nq,
                                                                   const int32_t* __restrict__ q_order, float* __restrict__ dist,
                                                                   int32_t* __restrict__ idx) {
  extern __shared__ __align__(16) float sbox[];          // [NC + NS][8]: boxes, then super-boxes
  const int b = blockIdx.y;
  const int NP = bvh_padded(n), NC = NP / BVH_CLUSTER, NS = bvh_supers(n);
  const uint8_t* base = bvh + (size_t)(sets_t == 1 ? 0 : b) * bvh_set_bytes(n);
  const float4* pts = reinterpret_cast<const float4*>(base);
  const float4* gbox = reinterpret_cast<const float4*>(base + (size_t)NP * 16);
  for (int i = threadIdx.x; i < (NC + NS) * 2; i += BVQ_THREADS) reinterpret_cast<float4*>(sbox)[i] = gbox[i];
  __syncthreads();
  const int slot = blockIdx.x * BVQ_THREADS + threadIdx.x;
  if (slot >= nq) return;
  const int qi = q_order ? q_order[slot] : slot;
  const float* qp = q + ((int64_t)(sets_q == 1 ? 0 : b) * nq + qi) * 3;
  const float qx = qp[0], qy = qp[1], qz = qp[2];

  auto lower = [&](int c) {
    const float4 lo = reinterpret_cast<const float4*>(sbox)[2 * c], hi = reinterpret_cast<const float4*>(sbox)[2 * c + 1];
    const float dx = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.f), dy = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.f),
                dz = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.f);
    // rounded down a little: the bound must never exceed the distance sqdist_ref reports for a point of the box
    return (dx * dx + dy * dy + dz * dz) * 0.999999f;
  };
  float best = INFINITY;
  int besti = 0x7fffffff;
  auto visit = [&](int c) {
    const float4* cp = pts + (size_t)c * BVH_CLUSTER;
#pragma unroll 8
    for (int j = 0; j < BVH_CLUSTER; ++j) {
      const float4 tp = __ldg(cp + j);
      const float d = sqdist_ref_bvh(tp.x, tp.y, tp.z, qx, qy, qz);
      const int ti = __float_as_int(tp.w);
      if (d < best || (d == best && ti < besti)) { best = d; besti = ti; }     // lowest original index on ties
    }
  };
  // pass 1: closest super-box -> its closest box -> a first candidate
  int s_first = 0, first = 0;
  float lb_first = INFINITY;
  for (int s = 0; s < NS; ++s) {
    const float lb = lower(NC + s);
    if (lb < lb_first) { lb_first = lb; s_first = s; }
  }
  lb_first = INFINITY;
  first = s_first * BVH_SUPER;
  for (int c = s_first * BVH_SUPER; c < min(NC, (s_first + 1) * BVH_SUPER); ++c) {
    const float lb = lower(c);
    if (lb < lb_first) { lb_first = lb; first = c; }
  }
  visit(first);
  // pass 2: everything that can still hold a point at most as far
  for (int s = 0; s < NS; ++s) {
    if (!(lower(NC + s) <= best)) continue;
    const int c1 = min(NC, (s + 1) * BVH_SUPER);
    for (int c = s * BVH_SUPER; c < c1; ++c) {
      if (c == first) continue;
      if (lower(c) <= best) visit(c);
    }
  }
  dist[(int64_t)b * nq + qi] = best;
  idx[(int64_t)b * nq + qi] = besti;
}

// Warp-cooperative variant: a warp owns 32 consecutive query slots and answers them ONE AT A TIME with all 32 lanes -- lane = super-box
// (<= 32 of them), then lane = box of an opened super-box (16), then lane = point of a visited cluster (32, one coalesced 512-byte read),
// an arg-min butterfly per visited cluster.  Same visiting rule and tie rule, hence the same (distance, index) as the per-thread kernel,
// but no divergence: the per-thread kernel ran with 16 of 32 lanes active (profiles/r1_ncu_nn_bvh.md).  Measured on B200: 2.9x SLOWER
// than the per-thread kernel (one dependent chain per warp instead of 16 independent ones) -- kept as a cross-check, not the default.
__global__ void __launch_bounds__(BVQ_THREADS) nn_bvh_query_warp_kernel(const uint8_t* __restrict__ bvh, int sets_t, int n,
                                                                        const float* __restrict__ q, int sets_q, int nq,
                                                                        const int32_t* __restrict__ q_order, float* __restrict__ dist,
                                                                        int32_t* __restrict__ idx) {
  extern __shared__ __align__(16) float sbox[];          // [NC + NS][8]
  const unsigned FULL = 0xffffffffu;
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int NP = bvh_padded(n), NC = NP / BVH_CLUSTER, NS = bvh_supers(n);      // NS <= 32 (n <= 16384)
  const uint8_t* base = bvh + (size_t)(sets_t == 1 ? 0 : b) * bvh_set_bytes(n);
  const float4* pts = reinterpret_cast<const float4*>(base);
  const float4* gbox = reinterpret_cast<const float4*>(base + (size_t)NP * 16);
  for (int i = threadIdx.x; i < (NC + NS) * 2; i += BVQ_THREADS) reinterpret_cast<float4*>(sbox)[i] = gbox[i];
  __syncthreads();
  const int slot = blockIdx.x * BVQ_THREADS + threadIdx.x;
  const bool live = slot < nq;
  const int qi = live ? (q_order ? q_order[slot] : slot) : 0;
  const float* qp = q + ((int64_t)(sets_q == 1 ? 0 : b) * nq + qi) * 3;
  const float myx = live ? qp[0] : 0.f, myy = live ? qp[1] : 0.f, myz = live ? qp[2] : 0.f;
  float my_best = INFINITY;
  int my_besti = 0x7fffffff;
  const unsigned live_mask = __ballot_sync(FULL, live);

  for (int s = 0; s < 32; ++s) {
    if (!((live_mask >> s) & 1u)) continue;                                   // warp-uniform
    const float qx = __shfl_sync(FULL, myx, s), qy = __shfl_sync(FULL, myy, s), qz = __shfl_sync(FULL, myz, s);
    auto lower = [&](int c) {
      const float4 lo = reinterpret_cast<const float4*>(sbox)[2 * c], hi = reinterpret_cast<const float4*>(sbox)[2 * c + 1];
      const float dx = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.f), dy = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.f),
                  dz = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.f);
      return (dx * dx + dy * dy + dz * dz) * 0.999999f;
    };
    float best = INFINITY;
    int besti = 0x7fffffff;
    auto visit = [&](int c) {                                                 // all lanes: one point each, then an arg-min butterfly
      const float4 tp = __ldg(pts + (size_t)c * BVH_CLUSTER + lane);
      float d = sqdist_ref_bvh(tp.x, tp.y, tp.z, qx, qy, qz);
      int ti = __float_as_int(tp.w);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float d2 = __shfl_xor_sync(FULL, d, o);
        const int t2 = __shfl_xor_sync(FULL, ti, o);
        if (d2 < d || (d2 == d && t2 < ti)) { d = d2; ti = t2; }
      }
      if (d < best || (d == best && ti < besti)) { best = d; besti = ti; }     // identical in every lane
    };
    // pass 1: closest super-box (lane = super-box), its closest box (lane = box), a first candidate
    const float slb = lane < NS ? lower(NC + lane) : INFINITY;
    float m = slb;
    int ms = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(FULL, m, o);
      const int s2 = __shfl_xor_sync(FULL, ms, o);
      if (m2 < m || (m2 == m && s2 < ms)) { m = m2; ms = s2; }
    }
    const int s_first = ms < NS ? ms : 0;
    {
      const int c = s_first * BVH_SUPER + lane;
      float bl = (lane < BVH_SUPER && c < NC) ? lower(c) : INFINITY;
      int bc = c;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float b2 = __shfl_xor_sync(FULL, bl, o);
        const int c2 = __shfl_xor_sync(FULL, bc, o);
        if (b2 < bl || (b2 == bl && c2 < bc)) { bl = b2; bc = c2; }
      }
      const int first = (bc < NC && bc >= s_first * BVH_SUPER && bc < (s_first + 1) * BVH_SUPER) ? bc : s_first * BVH_SUPER;
      visit(first);
      // pass 2: every super-box / box that can still hold a point at most as far
      for (int sb = 0; sb < NS; ++sb) {
        const float lbs = __shfl_sync(FULL, slb, sb);
        if (!(lbs <= best)) continue;                                         // warp-uniform
        const int cc = sb * BVH_SUPER + lane;
        const bool okc = lane < BVH_SUPER && cc < NC;
        const float lbc = okc ? lower(cc) : INFINITY;
        unsigned open = __ballot_sync(FULL, okc && lbc <= best && cc != first);
        while (open) {
          const int l = __ffs(open) - 1;
          open &= open - 1;
          const float lbl = __shfl_sync(FULL, lbc, l);
          if (lbl <= best) visit(sb * BVH_SUPER + l);                         // `best` may have shrunk since the ballot
        }
      }
    }
    if (lane == s) { my_best = best; my_besti = besti; }
  }
  if (live) {
    dist[(int64_t)b * nq + qi] = my_best;
    idx[(int64_t)b * nq + qi] = my_besti;
  }
}

}  // namespace zs

using namespace zs;

extern "C" size_t zs_nn_bvh_bytes(int sets, int n) {
  if (sets <= 0 || n <= 0) return 0;
  return (size_t)sets * bvh_set_bytes(n);
}

extern "C" int zs_nn_bvh_build(const float* pts, int sets, int n, void* bvh, void* stream) {
  ZS_REQUIRE(pts && bvh && sets > 0 && n > 0, "zs_nn_bvh_build: bad args");
  ZS_REQUIRE(n <= BVH_MAX_N, "zs_nn_bvh_build: at most %d points per set", BVH_MAX_N);
  ZS_REQUIRE((reinterpret_cast<uintptr_t>(bvh) & 15) == 0, "zs_nn_bvh_build: bvh buffer must be 16-byte aligned");
  const int smem = (int)sizeof(typename BvhSort::TempStorage);
  static thread_local bool configured = false;
  if (!configured) {
    ZS_CUDA_CALL(cudaFuncSetAttribute(nn_bvh_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  nn_bvh_build_kernel<<<sets, BVH_THREADS, smem, as_stream(stream)>>>(pts, n, reinterpret_cast<uint8_t*>(bvh));
  ZS_CUDA_CHECK_LAUNCH("zs_nn_bvh_build");
  return ZS_OK;
}

extern "C" int zs_nn_bvh_query(const void* bvh, int sets_t, int n, const float* q, int sets_q, int nq, int batch,
                               const int32_t* q_order, float* dist, int32_t* idx, int variant, void* stream) {
  ZS_REQUIRE(bvh && q && dist && idx && n > 0 && nq > 0 && batch > 0, "zs_nn_bvh_query: bad args");
  ZS_REQUIRE((sets_t == 1 || sets_t == batch) && (sets_q == 1 || sets_q == batch),
             "zs_nn_bvh_query: target / query set counts must be 1 (shared) or the batch size");
  ZS_REQUIRE(n <= BVH_MAX_N && batch <= 65535, "zs_nn_bvh_query: too many points per set or too large a batch");
  const int smem = (bvh_padded(n) / BVH_CLUSTER + bvh_supers(n)) * 32;
  dim3 grid((nq + BVQ_THREADS - 1) / BVQ_THREADS, batch);
  ZS_REQUIRE(variant == 0 || variant == 1, "zs_nn_bvh_query: variant must be 0 (thread per query) or 1 (warp-cooperative)");
  if (variant == 1)
    nn_bvh_query_warp_kernel<<<grid, BVQ_THREADS, smem, as_stream(stream)>>>(reinterpret_cast<const uint8_t*>(bvh), sets_t, n, q, sets_q,
                                                                            nq, q_order, dist, idx);
  else
    nn_bvh_query_kernel<<<grid, BVQ_THREADS, smem, as_stream(stream)>>>(reinterpret_cast<const uint8_t*>(bvh), sets_t, n, q, sets_q, nq,
                                                                       q_order, dist, idx);
  ZS_CUDA_CHECK_LAUNCH("zs_nn_bvh_query");
  return ZS_OK;
}
