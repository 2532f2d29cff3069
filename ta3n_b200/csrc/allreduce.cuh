// allreduce.cuh -- in-place mean all-reduce of the flat gradient bucket over NVLink / NVSwitch peer memory, written
// for the one collective of the path (the gradients of the used parameters, 13.9 MB at cfg2; replaces the
// reduce-to-GPU-0 of nn.DataParallel, main.py:79).
//
// The bucket lives in a SYMMETRIC allocation (same size on every rank of the node, peer-mapped; optionally also
// mapped through an NVSwitch multicast object).  Two-shot algorithm, one kernel, no host involvement:
//   barrier 0 : every rank's bucket is complete (flags written into the peers' flag arrays with release stores);
//   rank r owns slice r of the bucket: it reads that slice from every rank and sums it in rank order --
//               multimem.ld_reduce (the switch adds the 8 copies in flight, NVLS) when a multicast mapping exists,
//               peer loads otherwise -- scales by 1/world and writes the result into every rank's bucket
//               (multimem.st / peer stores);
//   barrier 1 : every slice has landed everywhere -> the kernel may complete, the optimizer kernel reads the mean.
// Every slice is summed by exactly one rank and broadcast, so all ranks hold bit-identical gradients; the order of the
// sum is fixed (rank order / switch order), so reruns are bit-identical too.  The kernel is CUDA-graph capturable
// (plain device pointers), which NCCL collectives were not in practice (round 1: process-group teardown hang).
#pragma once

#include "common.cuh"

namespace ta3n {

constexpr int kArMaxWorld = 16;
constexpr int kArBlocks = 128;
constexpr int kArThreads = 512;
constexpr int kArUnroll = 4;          // independent 16 B reductions a thread keeps in flight (the loop is latency-bound)

struct ArPeers {
  float* buf[kArMaxWorld];          // peer mappings of the bucket ([rank] = local)
  unsigned* flags[kArMaxWorld];     // peer mappings of the flag arrays: [2][kArBlocks][world] uint32 each
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, const float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all blocks `b` of all ranks meet: slot = (phase, b); one flag per source rank.
// Only the `world` signalling threads of a block fence (a system-scope fence by each of the 65 k threads of the grid
// cost ~15 us per barrier: measured, tools/allreduce_probe.py), cumulatively for the block through the CTA barrier;
// polls are relaxed loads (an acquire load is a load + L1 invalidate on this part) with one fence after the wait.
__device__ __forceinline__ void ar_barrier(const ArPeers& P, int rank, int world, int phase, unsigned seq) {
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int p = threadIdx.x;
    const size_t slot = ((size_t)phase * kArBlocks + blockIdx.x) * world;
    __threadfence_system();                       // this block's writes, then the flag
    st_release_sys(P.flags[p] + slot + rank, seq);
    unsigned spins = 0;
    while ((int)(ld_relaxed_sys(P.flags[rank] + slot + p) - seq) < 0) {
      if (++spins > (1u << 26)) __trap();         // a missing peer becomes an error, not a hung GPU
    }
    __threadfence_system();                       // acquire side
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kArThreads, 1)
allreduce_mean_kernel(const __grid_constant__ ArPeers P, float* __restrict__ mc, const unsigned long long* __restrict__ seq_dev,
                      const int rank, const int world, const size_t n4, const float scale) {
  pdl_wait();
  const unsigned seq = (unsigned)seq_dev[0];
  ar_barrier(P, rank, world, 0, seq);
  const size_t s0 = n4 * (size_t)rank / world, s1 = n4 * (size_t)(rank + 1) / world;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (mc != nullptr) {
    for (size_t i0 = s0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < s1; i0 += stride * kArUnroll) {
      float4 v[kArUnroll];
#pragma unroll
      for (int u = 0; u < kArUnroll; ++u) {
        const size_t i = i0 + (size_t)u * stride;
        if (i < s1) v[u] = multimem_ld_reduce_add(mc + 4 * i);
      }
#pragma unroll
      for (int u = 0; u < kArUnroll; ++u) {
        const size_t i = i0 + (size_t)u * stride;
        if (i < s1) {
          v[u].x *= scale;
          v[u].y *= scale;
          v[u].z *= scale;
          v[u].w *= scale;
          multimem_st(mc + 4 * i, v[u]);
        }
      }
    }
  } else {
    for (size_t i0 = s0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < s1; i0 += stride * 2) {
      float4 acc[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const size_t i = i0 + (size_t)u * stride;
        acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < s1) {
#pragma unroll 8
          for (int p = 0; p < world; ++p) {       // fixed rank order; `world` independent loads in flight
            const float4 x = __ldcg(reinterpret_cast<const float4*>(P.buf[p]) + i);
            acc[u].x += x.x;
            acc[u].y += x.y;
            acc[u].z += x.z;
            acc[u].w += x.w;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const size_t i = i0 + (size_t)u * stride;
        if (i < s1) {
          acc[u].x *= scale;
          acc[u].y *= scale;
          acc[u].z *= scale;
          acc[u].w *= scale;
#pragma unroll 8
          for (int p = 0; p < world; ++p) reinterpret_cast<float4*>(P.buf[p])[i] = acc[u];
        }
      }
    }
  }
  ar_barrier(P, rank, world, 1, seq);
}

}  // namespace ta3n
