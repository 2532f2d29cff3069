// gemm_tcgen05.cuh -- 5th-generation tensor-core engine for the segmented grouped GEMM.
//
// One CTA per 128 x 128 output tile (2 CTAs resident per SM), warp-specialised:
//   warp 0     : TMA producer  -- cp.async.bulk.tensor.2d into a 3-stage ring of 128B-swizzled tiles
//   warp 1     : TMEM allocator + single-thread tcgen05.mma.kind::tf32 issuer (fp32 accum in TMEM)
//   warps 2..5 : epilogue      -- tcgen05.ld (32 lanes x 32 columns: one accumulator row per lane),
//                                 fused epilogue (apply_epilogue) in registers, 16 B stores covering
//                                 each 128 B line of the output row completely
// Operands stay fp32 in HBM: kind::tf32 reads the fp32 bit patterns directly (10-bit mantissa,
// fp32 range), so there is no conversion pass and no second copy of any tensor.
// The "gather" of TRN frame tuples, the source/target split and the per-frame dgrad are all
// expressed as TMA coordinates / tensor maps per K-segment: nothing is materialised.
//
// Both operand majors are supported through the UMMA smem descriptors:
//   K-major  (A(m,k)=A[m*ld+k]) : one TMA box {32 k, 128 rows}, SWIZZLE_128B (16 B chunks);
//                                 desc layout SW128, SBO=1024 B, +32 B per K=8 step
//   MN-major (A(m,k)=A[k*ld+m]) : four TMA boxes {32 m, 32 k}, SWIZZLE_128B_ATOM_32B -- the only smem
//                                 layout tcgen05 accepts for MN-major 32-bit operands (swizzle atom =
//                                 4 k-rows x 128 B); desc layout SW128_BASE32B, LBO=4096 B (next 32 m),
//                                 SBO=512 B (next 4 k-rows), +1024 B per K=8 step
#pragma once

#include <cuda.h>
#include <stdlib.h>

#include <algorithm>
#include <functional>
#include <map>
#include <tuple>

#include "seg_gemm.cuh"

namespace ta3n {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32;
// warps 0,1 = TMA / MMA; then 4 epilogue warps (3-stage variant: two CTAs share an SM's registers) or 8
// (6-stage variant, one CTA per SM: two warps per TMEM lane quarter, each finishing half of the columns)
__host__ __device__ constexpr int tc_threads(int stages) { return stages <= 3 ? 192 : 320; }
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;   // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;   // 16 KB
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
// Two pipeline depths: 3 stages (97 KB, two CTAs per SM: one CTA's epilogue overlaps the other's main
// loop) for grids beyond a wave; 6 stages (193 KB, one CTA per SM, twice the bytes in flight per CTA)
// for the sub-wave grids of this workload, where a CTA is alone on its SM and TMA latency-bound.
constexpr int tc_smem_bytes(int stages) { return stages * TC_STAGE_BYTES + 1024; }   // + 1024 B alignment slack
// X3 ("tf32x3", experimental engine 2): fp32-grade products from three tf32 MMAs per K step.  A stage then also
// holds the low-order halves of both operand tiles (64 KB), three stages, one CTA per SM, 8 helper/epilogue warps.
constexpr int TC_X3_STAGES = 3;
constexpr int tc_x3_smem_bytes() { return TC_X3_STAGES * 2 * TC_STAGE_BYTES + 1024; }
constexpr int TC_TMEM_COLS = 128;
#ifndef TA3N_MAX_MAPS
#define TA3N_MAX_MAPS 64
#endif
constexpr int kMaxMaps = TA3N_MAX_MAPS;

struct alignas(64) TcMaps {
  CUtensorMap m[kMaxMaps];
};
struct TcSegMaps {
  unsigned char a[kMaxSegs];
  unsigned char b[kMaxSegs];
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// L2 prefetch of a tile (no shared-memory destination, no barrier): warms the line ahead of the real load
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: lane i of the warp receives row (lane base + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 |
//   [61,64) layout: 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7u) << 61;
  return d;
}
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t tile_base, int kstep) {
  return umma_desc(tile_base + kstep * 32, 16, 1024, 2);
}
__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t tile_base, int kstep) {
  return umma_desc(tile_base + kstep * 1024, 4096, 512, 1);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a,b=TF32 [7,10)=[10,13)=2,
// a_major bit 15, b_major bit 16 (1 = MN-major), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(bool a_kmaj, bool b_kmaj, int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_kmaj ? 0u : 1u) << 15) | ((b_kmaj ? 0u : 1u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// a - trunc_tf32(a): what the tensor core drops when it reads raw fp32 bits (0 for inf / nan, which the leading
// product already propagates)
__device__ __forceinline__ float tf32_low(float a) {
  const uint32_t u = __float_as_uint(a);
  return (u & 0x7f800000u) == 0x7f800000u ? 0.f : a - __uint_as_float(u & 0xffffe000u);
}

// ---- the kernel -------------------------------------------------------------------------------------
// FIXUP (experimental, TA3N_FIXUP_SPLITK=1): groups with fix_slot >= 0 fold their split-K partials inside this
// kernel instead of a reduce pass.  Splits 0 .. ksplit-2 of a tile write their raw accumulators to `partial` and
// bump the tile's arrival counter (release); the LAST split acquires the counter after its own K range, adds the
// partials in a fixed order in its register epilogue, runs the fused epilogue and resets the counter for the next
// launch.  Nobody but the last split ever waits, and the host only plans such launches when every CTA is resident
// at once (<= 256 CTAs at two per SM), so the wait cannot starve the blocks it waits for; a bounded spin turns
// any violation of that assumption into a trap instead of a hang.
//
// X3 (experimental engine "tf32x3"): the tensor maps deliver RAW fp32 bits; the tensor core reads only the top 19
// bits of an operand, i.e. hi = trunc_tf32(a).  The helper warps (idle during the main loop otherwise) compute
// lo = a - hi (exact in fp32) for both tiles of a stage into a second pair of buffers, and the MMA thread issues
// A*B + A*B_lo + A_lo*B per K step: relative error ~3 * 2^-20 per product instead of 2^-11, at unchanged operand
// traffic -- the main loop is bound by the fill bandwidth, the tensor pipe is 26 % busy with one MMA per step.
template <bool A_KMAJ, bool B_KMAJ, int TC_STAGES, bool FIXUP, bool X3 = false>
__global__ void __launch_bounds__(X3 ? 320 : tc_threads(TC_STAGES), (X3 || TC_STAGES > 3) ? 1 : 2)
seg_gemm_tc_kernel(const __grid_constant__ GemmTable tab, const __grid_constant__ TcMaps maps,
                   const __grid_constant__ TcSegMaps segmaps, const int dbg, int* __restrict__ fix_flags) {
  static_assert(!X3 || (TC_STAGES == TC_X3_STAGES && !FIXUP), "tf32x3 variant: 3 stages, no fix-up");
  constexpr int kThreads = X3 ? 320 : tc_threads(TC_STAGES);
  constexpr int kStageStride = X3 ? 2 * TC_STAGE_BYTES : TC_STAGE_BYTES;
  extern __shared__ uint8_t tc_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TC_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TC_STAGES];
  __shared__ __align__(8) uint64_t split_bar[TC_STAGES];     // X3: low-order tiles of the stage are ready
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- tile decode (same scheme as the SIMT engine, 128x128 tiles); group + segments staged in smem ----
  __shared__ TileCtx ctx;
  // Groups arrive sorted by K (longest first).  CTAs beyond the first wave of 148 take tiles from the END
  // of the list, so the SM that received the longest tile gets the shortest one as its second resident CTA.
  int tile = blockIdx.x;
  if (tile >= 148) tile = tab.total_tiles - 1 - (tile - 148);
  load_tile_ctx(tab, tile, &ctx, segmaps.a, segmaps.b);
  const Group& g = ctx.g;
  int local = tile - g.tile_begin;
  const int per_split = g.tiles_m * g.tiles_n;
  const int split = local / per_split;
  local -= split * per_split;
  const int m0 = (local / g.tiles_n) * TC_BM;
  const int n0 = (local % g.tiles_n) * TC_BN;

  int total_chunks = 0;
  for (int s = 0; s < g.seg_count; ++s) total_chunks += (ctx.seg[s].len + TC_BK - 1) / TC_BK;
  const int cps = (total_chunks + g.ksplit - 1) / g.ksplit;
  const int c_begin = split * cps;
  const int c_end = min(total_chunks, c_begin + cps);
  int n_iter = max(0, c_end - c_begin);
  if (dbg & 4) n_iter = 0;       // debug: no TMA / MMA
  if (dbg & 8) return;           // debug: nothing at all

  // ---- one-time setup ----
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&split_bar[s], kThreads / 32 - 2);      // one arrival per helper warp (X3 only)
    }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1 && !(dbg & 2)) tmem_alloc(&tmem_slot, TC_TMEM_COLS);
  if (warp == 0 && (dbg & (1 << 24))) {
    // Experimental (TA3N_DESC_PREFETCH=1, default off): the tensor maps are kernel parameters, so their
    // descriptors can be pulled into the TMA unit's cache here, under the previous kernel's tail, instead of
    // stalling the first load of every segment.
    for (int sgi = lane; sgi < g.seg_count; sgi += 32) {
      tma_prefetch_desc(&maps.m[ctx.seg[sgi].amap]);
      tma_prefetch_desc(&maps.m[ctx.seg[sgi].bmap]);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // Everything above touched only kernel parameters, shared memory and TMEM: it overlaps the previous
  // kernel of the stream.  From here on operands produced by that kernel are read.
  pdl_wait();
  if (dbg & 16) {                // debug: setup + teardown only
    __syncthreads();
    if (warp == 1 && !(dbg & 2)) tmem_dealloc(tmem_base, TC_TMEM_COLS);
    return;
  }

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0 && n_iter > 0) {
      int seg = 0, k0 = 0;
      {
        int skip = c_begin;
        while (seg < g.seg_count) {
          const int nch = (ctx.seg[seg].len + TC_BK - 1) / TC_BK;
          if (skip < nch) {
            k0 = skip * TC_BK;
            break;
          }
          skip -= nch;
          ++seg;
        }
      }
      // Experimental (TA3N_L2_PREFETCH=<slabs>, default 0 = off): TMA prefetch into L2 `pf` slabs ahead of the
      // loads.  The first GEMMs of a step read inputs and weights that are cold in L2 (HBM latency ~2x an L2 hit)
      // with only TC_STAGES slabs in flight per SM; an L2 prefetch needs no shared memory, so it can run far ahead.
      const int pf = (dbg >> 16) & 0xff;
      int pseg = seg, pk0 = k0, pleft = 0;
      if (pf > 0) {
        pleft = n_iter;
        for (int a = 0; a < pf && pleft > 0; ++a) {      // skip the first pf slabs: the real loads fetch those
          --pleft;
          pk0 += TC_BK;
          if (pk0 >= ctx.seg[pseg].len) {
            ++pseg;
            pk0 = 0;
          }
        }
      }
      for (int it = 0; it < n_iter; ++it) {
        if (pleft > 0) {
          const CUtensorMap* pa = &maps.m[ctx.seg[pseg].amap];
          const CUtensorMap* pb = &maps.m[ctx.seg[pseg].bmap];
          if (A_KMAJ) {
            tma_prefetch_2d(pa, pk0, m0);
          } else if (tab.pad_ & 1) {
            tma_prefetch_3d(pa, 0, pk0, m0 >> 5);
          }
          if (B_KMAJ) {
            tma_prefetch_2d(pb, pk0, n0);
          } else if (tab.pad_ & 2) {
            tma_prefetch_3d(pb, 0, pk0, n0 >> 5);
          }
          --pleft;
          pk0 += TC_BK;
          if (pk0 >= ctx.seg[pseg].len) {
            ++pseg;
            pk0 = 0;
          }
        }
        const int stage = it % TC_STAGES;
        const uint32_t phase = (uint32_t)(it / TC_STAGES) & 1u;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_expect_tx(&full_bar[stage], TC_STAGE_BYTES);
        uint8_t* sA = smem + stage * kStageStride;
        uint8_t* sB = sA + TC_A_BYTES;
        const CUtensorMap* ma = &maps.m[ctx.seg[seg].amap];
        const CUtensorMap* mb = &maps.m[ctx.seg[seg].bmap];
        // MN-major tiles are four [32 k-rows][32 floats] slabs, one per group of 32 m (or n).  When the operand's
        // MN extent is a multiple of 32 a rank-3 tensor map {32 floats, k rows, groups of 32} fetches all four with
        // ONE instruction (the lone producer thread is issue-bound: ~50 cycles per TMA, 8 per chunk otherwise).
        if (A_KMAJ) {
          tma_load_2d(sA, ma, &full_bar[stage], k0, m0);
        } else if (tab.pad_ & 1) {
          tma_load_3d(sA, ma, &full_bar[stage], 0, k0, m0 >> 5);
        } else {
#pragma unroll
          for (int q = 0; q < TC_BM / 32; ++q) tma_load_2d(sA + q * 4096, ma, &full_bar[stage], m0 + 32 * q, k0);
        }
        if (B_KMAJ) {
          tma_load_2d(sB, mb, &full_bar[stage], k0, n0);
        } else if (tab.pad_ & 2) {
          tma_load_3d(sB, mb, &full_bar[stage], 0, k0, n0 >> 5);
        } else {
#pragma unroll
          for (int q = 0; q < TC_BN / 32; ++q) tma_load_2d(sB + q * 4096, mb, &full_bar[stage], n0 + 32 * q, k0);
        }
        k0 += TC_BK;
        if (k0 >= ctx.seg[seg].len) {
          ++seg;
          k0 = 0;
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (one thread) ===========================
    if (lane == 0 && n_iter > 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(A_KMAJ, B_KMAJ, TC_BM, TC_BN);
      for (int it = 0; it < n_iter; ++it) {
        const int stage = it % TC_STAGES;
        const uint32_t phase = (uint32_t)(it / TC_STAGES) & 1u;
        mbar_wait(X3 ? &split_bar[stage] : &full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + stage * kStageStride);
        const uint32_t b_base = a_base + TC_A_BYTES;
#pragma unroll
        for (int ks = 0; ks < TC_BK / 8; ++ks) {
          if ((dbg & 512) && ks > 0) break;   // debug: 1 of 4 MMAs (timing experiments only, wrong results)
          const uint64_t adesc = A_KMAJ ? umma_desc_kmajor(a_base, ks) : umma_desc_mnmajor(a_base, ks);
          const uint64_t bdesc = B_KMAJ ? umma_desc_kmajor(b_base, ks) : umma_desc_mnmajor(b_base, ks);
          if (X3) {
            // small terms first, then the leading product; the low-order tiles sit TC_STAGE_BYTES above their tiles
            const uint32_t al = a_base + TC_STAGE_BYTES, bl = b_base + TC_STAGE_BYTES;
            const uint64_t aldesc = A_KMAJ ? umma_desc_kmajor(al, ks) : umma_desc_mnmajor(al, ks);
            const uint64_t bldesc = B_KMAJ ? umma_desc_kmajor(bl, ks) : umma_desc_mnmajor(bl, ks);
            umma_tf32(tmem_base, adesc, bldesc, idesc, (it > 0 || ks > 0) ? 1u : 0u);
            umma_tf32(tmem_base, aldesc, bdesc, idesc, 1u);
            umma_tf32(tmem_base, adesc, bdesc, idesc, 1u);
          } else {
            umma_tf32(tmem_base, adesc, bdesc, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);   // releases the smem slot once these MMAs have read it
      }
      umma_commit(&tmem_full_bar);        // accumulator complete
    }
  } else {
    // =========================== epilogue (4 warps, 32 TMEM lanes each) ===========================
    // Lane i of warp lq owns accumulator row 32*lq + i; tcgen05.ld hands it 32 consecutive columns, which
    // it finishes (fused epilogue) and writes as 8 x 16 B stores: every 128 B line of C is written whole.
    if (X3 && n_iter > 0) {
      // ---- tf32x3 helper role: lo = a - trunc_tf32(a) for the A and B tiles of every stage (elementwise, so the
      // swizzled layout is irrelevant), written TC_STAGE_BYTES above the tiles with the same layout
      constexpr int kHelpers = kThreads - 64;
      const int ht = threadIdx.x - 64;
      for (int it = 0; it < n_iter; ++it) {
        const int stage = it % TC_STAGES;
        const uint32_t phase = (uint32_t)(it / TC_STAGES) & 1u;
        mbar_wait(&full_bar[stage], phase);
        float4* tile = reinterpret_cast<float4*>(smem + stage * kStageStride);
        float4* low = reinterpret_cast<float4*>(smem + stage * kStageStride + TC_STAGE_BYTES);
#pragma unroll
        for (int i = ht; i < TC_STAGE_BYTES / 16; i += kHelpers) {
          const float4 a = tile[i];
          low[i] = make_float4(tf32_low(a.x), tf32_low(a.y), tf32_low(a.z), tf32_low(a.w));
        }
        fence_proxy_async();              // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[stage]);
      }
    }
    const int lq = warp & 3;              // TMEM lane quarter this warp may access
    const Group e = ctx.g;                // register copy: no reloads behind the global stores
    const int m = m0 + lq * 32 + lane;
    const bool fix = FIXUP && e.ksplit > 1 && e.fix_slot >= 0;
    const bool owner = fix && split == e.ksplit - 1;
    const bool split_out = e.ksplit > 1 && !owner;
    float* const obase = split_out ? e.partial + (size_t)split * e.M * e.N : e.C;
    const int ldo = split_out ? e.N : e.ldc;
    int* const flag = fix ? fix_flags + e.fix_slot + local : nullptr;
    if (n_iter > 0) {
      mbar_wait(&tmem_full_bar, 0);
      tc_fence_after();
    }
    if (FIXUP && owner) {
      // wait until every other split of this tile has stored its partial (they are resident and never wait)
      if (lane == 0) {
        int seen = 0;
        for (unsigned spin = 0; ; ++spin) {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
          if (seen >= e.ksplit - 1) break;
          if (spin > (1u << 22)) __trap();      // seconds: a protocol error becomes a trap, not a hung device
          __nanosleep(128);
        }
      }
      __syncwarp();
    }
    constexpr int kEpiWarps = kThreads / 32 - 2;
    constexpr int kColChunks = (TC_BN / 32) * 4 / kEpiWarps;      // column chunks of 32 per warp: 4 or 2
    const int c0 = ((warp - 2) / 4) * kColChunks;
#pragma unroll 1
    for (int c = c0; c < c0 + kColChunks; ++c) {
      float v[32];
      if (n_iter > 0 && !(dbg & 128)) {
        tmem_ld_32x32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(c * 32), v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      const int nb = n0 + c * 32;
      if (m < e.M && nb < e.N && !(dbg & 1) && (!(dbg & 32) || lane == 0)) {
        float* orow = obase + (size_t)m * ldo + nb;
        if (dbg & 64) {
          orow[0] = v[0];
          continue;
        }
        if (FIXUP && owner) {      // raw partial sums of the other splits, fixed order, straight from L2
          const int nvalid = min(32, e.N - nb);
          for (int sp = 0; sp < e.ksplit - 1; ++sp) {
            const float* pr = e.partial + (size_t)sp * e.M * e.N + (size_t)m * e.N + nb;
            if (nvalid >= 32 && (reinterpret_cast<uintptr_t>(pr) & 15u) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 t = __ldcg(reinterpret_cast<const float4*>(pr + j));
                v[j] += t.x;
                v[j + 1] += t.y;
                v[j + 2] += t.z;
                v[j + 3] += t.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nvalid) v[j] += __ldcg(pr + j);
            }
          }
        }
        if (!split_out && !(dbg & 256)) {
          const int nvalid = min(32, e.N - nb);
          TA3N_EPI_DISPATCH(e.flags, { epilogue_row32<EPI_F>(e, m, nb, nvalid, v); })
        }
        if (nb + 32 <= e.N && ((reinterpret_cast<uintptr_t>(orow) & 15u) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(orow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < e.N) orow[j] = v[j];
        }
      }
    }
    if (FIXUP && fix) {
      // all epilogue warps have issued their stores -> one thread publishes (or, for the owner, re-arms) the counter
      __threadfence();
      asm volatile("bar.sync 1, %0;" ::"n"((kThreads / 32 - 2) * 32) : "memory");
      if (warp == 2 && lane == 0) {
        if (owner)
          *reinterpret_cast<volatile int*>(flag) = 0;
        else
          asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(flag) : "memory");
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1 && !(dbg & 2)) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled encode_fn() {
  static PFN_encodeTiled fn = []() -> PFN_encodeTiled {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<PFN_encodeTiled>(p);
  }();
  return fn;
}

struct MapKey {
  const void* ptr;
  long inner, outer, ld;
  int box_inner, box_outer;
  int atom32;   // 1: SWIZZLE_128B_ATOM_32B (MN-major operands), 0: SWIZZLE_128B
  int rank3;    // 1: MN-major operand as {32 floats, outer rows, inner/32 groups}, box {32, box_outer, 4}
  int raw;      // 1: FLOAT32 map (bits as stored; tf32x3 engine), 0: TFLOAT32 (TMA rounds to nearest tf32)
  bool operator<(const MapKey& o) const {
    return std::tie(ptr, inner, outer, ld, box_inner, box_outer, atom32, rank3, raw) <
           std::tie(o.ptr, o.inner, o.outer, o.ld, o.box_inner, o.box_outer, o.atom32, o.rank3, o.raw);
  }
};

// fp32 2-D row-major tensor [outer, inner] with row pitch ld floats; 128B swizzle; OOB reads give zeros
inline int encode_map(const MapKey& k, CUtensorMap* out) {
  static thread_local std::map<MapKey, CUtensorMap> cache;
  auto it = cache.find(k);
  if (it != cache.end()) {
    *out = it->second;
    return TA3N_OK;
  }
  PFN_encodeTiled fn = encode_fn();
  if (!fn) return fail(TA3N_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)k.inner, (cuuint64_t)k.outer, 1};
  cuuint64_t strides[2] = {(cuuint64_t)k.ld * sizeof(float), 128};
  cuuint32_t box[3] = {(cuuint32_t)k.box_inner, (cuuint32_t)k.box_outer, 4};
  cuuint32_t estr[3] = {1, 1, 1};
  if (k.rank3) {
    dims[0] = 32;
    dims[2] = (cuuint64_t)(k.inner / 32);
  }
  // TFLOAT32: the TMA unit rounds fp32 -> tf32 (round to nearest) while filling shared memory, so the
  // tensor core never sees the truncation bias (-2^-11 relative per operand) it would apply to raw fp32
  // bit patterns.  TA3N_TMA_RAW_FP32=1 switches to the raw copy (kept to measure the difference).
  static const bool raw = []() {
    const char* e = getenv("TA3N_TMA_RAW_FP32");
    return e && e[0] == '1';
  }();
  CUresult r = fn(out, (raw || k.raw) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, k.rank3 ? 3 : 2,
                  const_cast<void*>(k.ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  k.atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TA3N_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  if (cache.size() > 4096) cache.clear();
  cache[k] = *out;
  return TA3N_OK;
}

inline bool tc_operand_ok(const float* p, int ld) {
  return p != nullptr && (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && (ld % 4) == 0 && ld > 0;
}

// Is this group worth / able to run on the tensor-core engine?
inline bool tc_group_ok(const GemmPlan& plan, const Group& g) {
  if (plan.load_flags != 0) return false;                    // ReLU-on-load needs a register pass
  if ((long)g.M * g.N < 64L * 64L) return false;             // tiny heads stay on the SIMT engine
  // Experimental (TA3N_SIMT_MAX_MNK=<M*N*K>, default 0 = off): a tcgen05 launch costs ~7 us before its first MMA
  // retires (ncu: 8-slab launches), more than an fp32 SIMT pass over a 512x256x256 discriminator layer.
  static const double simt_below = []() {
    const char* e = getenv("TA3N_SIMT_MAX_MNK");
    return e ? atof(e) : 0.0;
  }();
  if (simt_below > 0.0 && (double)g.M * g.N * (double)plan.k_total(g) <= simt_below) return false;
  if (g.seg_count > kMaxSegs) return false;
  for (int i = 0; i < g.seg_count; ++i) {
    const Seg& s = plan.segs[g.seg_begin + i];
    if (!tc_operand_ok(s.A, s.lda) || !tc_operand_ok(s.B, s.ldb) || s.len <= 0) return false;
  }
  return true;
}

template <bool A_KMAJ, bool B_KMAJ, int STAGES, bool FIXUP>
inline int tc_launch_stages(const GemmTable& tab, const TcMaps& maps, const TcSegMaps& sm, cudaStream_t stream,
                            const char* label, int* fix_flags) {
  static bool configured = false;
  if (!configured) {
    TA3N_CUDA(cudaFuncSetAttribute(seg_gemm_tc_kernel<A_KMAJ, B_KMAJ, STAGES, FIXUP>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(STAGES)));
    configured = true;
  }
  static const int dbg = []() {
    const char* e = getenv("TA3N_TC_DEBUG");
    const char* p = getenv("TA3N_L2_PREFETCH");          // experimental: L2 prefetch distance in K slabs (0 = off)
    int pf = p ? atoi(p) : 0;
    pf = pf < 0 ? 0 : (pf > 255 ? 255 : pf);
    const char* d = getenv("TA3N_DESC_PREFETCH");        // experimental: prefetch tensor-map descriptors in the prologue
    return ((e ? atoi(e) : 0) & 0xffff) | (pf << 16) | ((d && d[0] == '1') ? (1 << 24) : 0);
  }();
  pre_launch(label, stream);
  launch_kernel(seg_gemm_tc_kernel<A_KMAJ, B_KMAJ, STAGES, FIXUP>, tab.total_tiles, tc_threads(STAGES),
                tc_smem_bytes(STAGES), stream, tab, maps, sm, dbg, fix_flags);
  return after_launch();
}

// tf32x3 engine: 3 stages of 64 KB, 320 threads, one CTA per SM
template <bool A_KMAJ, bool B_KMAJ>
inline int tc_launch_x3(const GemmTable& tab, const TcMaps& maps, const TcSegMaps& sm, cudaStream_t stream,
                        const char* label) {
  auto kernel = seg_gemm_tc_kernel<A_KMAJ, B_KMAJ, TC_X3_STAGES, false, true>;
  static bool configured = false;
  if (!configured) {
    TA3N_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_x3_smem_bytes()));
    configured = true;
  }
  pre_launch(label, stream);
  launch_kernel(kernel, tab.total_tiles, 320, tc_x3_smem_bytes(), stream, tab, maps, sm, 0, static_cast<int*>(nullptr));
  return after_launch();
}

template <bool A_KMAJ, bool B_KMAJ>
inline int tc_launch_one(const GemmTable& tab, const TcMaps& maps, const TcSegMaps& sm, cudaStream_t stream,
                         const char* label, int* fix_flags, bool x3 = false) {
  if (x3) return tc_launch_x3<A_KMAJ, B_KMAJ>(tab, maps, sm, stream, label);
  static const int force = []() {
    const char* e = getenv("TA3N_TC_STAGES");
    return e ? atoi(e) : 0;
  }();
  const bool deep = force ? force > 3 : tab.total_tiles <= 148;
  if (fix_flags != nullptr)
    return deep ? tc_launch_stages<A_KMAJ, B_KMAJ, 6, true>(tab, maps, sm, stream, label, fix_flags)
                : tc_launch_stages<A_KMAJ, B_KMAJ, 3, true>(tab, maps, sm, stream, label, fix_flags);
  return deep ? tc_launch_stages<A_KMAJ, B_KMAJ, 6, false>(tab, maps, sm, stream, label, nullptr)
              : tc_launch_stages<A_KMAJ, B_KMAJ, 3, false>(tab, maps, sm, stream, label, nullptr);
}

// ---- experimental: in-kernel split-K fix-up (TA3N_FIXUP_SPLITK=1) -----------------------------------------
inline bool fixup_enabled() {
  static const bool on = []() {
    const char* e = getenv("TA3N_FIXUP_SPLITK");
    return e && e[0] == '1';
  }();
  return on;
}

// Arrival counters: one int per split output tile, zero between launches (the owner re-arms its counter).  The
// buffer belongs to the library (one per device, allocated on first use OUTSIDE stream capture -- during capture
// an unallocated buffer simply disables the fix-up for that call); launches take consecutive slots from a ring.
constexpr int kFixFlagSlots = 1 << 16;
inline int* fix_flags_take(int n, cudaStream_t stream, int* slot) {
  struct Dev {
    int* base = nullptr;
    unsigned next = 0;
  };
  static std::mutex mu;
  static std::map<int, Dev> devs;
  if (n <= 0 || n > kFixFlagSlots / 4) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  Dev& d = devs[dev];
  if (!d.base) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return nullptr;
    int* p = nullptr;
    if (cudaMalloc(&p, sizeof(int) * kFixFlagSlots) != cudaSuccess) return nullptr;
    if (cudaMemset(p, 0, sizeof(int) * kFixFlagSlots) != cudaSuccess) {
      cudaFree(p);
      return nullptr;
    }
    d.base = p;
  }
  if (d.next + (unsigned)n > (unsigned)kFixFlagSlots) d.next = 0;
  *slot = (int)d.next;
  d.next += (unsigned)n;
  return d.base;
}

// Modelled makespan (in 32-wide K slabs) of a launch: CTAs in launch order, each to the least loaded of 148 SMs;
// an SM's fill bandwidth is shared by its resident CTAs, so loads add up.
inline long tc_makespan(const std::vector<long>& cta_slabs) {
  std::vector<long> sm(148, 0);
  std::vector<long> sorted = cta_slabs;
  std::sort(sorted.begin(), sorted.end(), std::greater<long>());
  for (long c : sorted) *std::min_element(sm.begin(), sm.end()) += c;
  return *std::max_element(sm.begin(), sm.end());
}

// Choose per-group split factors that shorten the modelled makespan of a launch whose tiles are few and uneven
// (the critical tile of the forward batch is 80 slabs against 36 per SM on average).  Pure host logic: `tiles[i]`
// output tiles of `slabs[i]` K slabs each -> split factor per group (all 1 when splitting is not worth it).
inline std::vector<int> balance_split_factors(const std::vector<long>& tiles, const std::vector<long>& slabs) {
  const int ng = (int)tiles.size();
  long total = 0;
  for (int i = 0; i < ng; ++i) total += tiles[i] * slabs[i];
  auto model = [&](const std::vector<int>& ks) {
    std::vector<long> ctas;
    for (int i = 0; i < ng; ++i)
      for (long t = 0; t < tiles[i] * ks[i]; ++t) ctas.push_back((slabs[i] + ks[i] - 1) / ks[i] + 2);   // +2: fix-up
    return tc_makespan(ctas);
  };
  std::vector<int> ones(ng, 1), best(ng, 1);
  const long before = model(ones);
  long best_span = before;
  const long target = std::max<long>(12, (long)(1.15 * (double)total / 148.0));
  for (int bump = 0; bump <= 1; ++bump) {          // the even-load split, and one step finer (80 tiles: 3 beats 2)
    std::vector<int> ks(ng, 1);
    long n_ctas = 0;
    for (int i = 0; i < ng; ++i) {
      int want = (int)((slabs[i] + target - 1) / target);
      if (want > 1 || bump) want += bump;
      want = std::min(want, 4);
      while (want > 1 && slabs[i] / want < 8) --want;
      ks[i] = std::max(want, 1);
      n_ctas += tiles[i] * ks[i];
    }
    if (n_ctas > 256) continue;                                // every CTA resident at once, with headroom
    const long span = model(ks);
    if (span < best_span) {
      best_span = span;
      best = ks;
    }
  }
  if (best_span * 100 > before * 85) return ones;              // needs >= 15 % shorter critical path
  return best;
}

// Only with the fix-up: a separate reduce pass costs more than the balance saves.
inline void plan_balance_splitk(GemmPlan& plan, Arena* arena) {
  if (!arena) return;
  const int ng = (int)plan.groups.size();
  std::vector<long> tiles(ng), slabs(ng);
  for (int i = 0; i < ng; ++i) {
    const Group& g = plan.groups[i];
    if (g.ksplit > 1) return;   // already planned
    tiles[i] = (long)((g.M + TC_BM - 1) / TC_BM) * ((g.N + TC_BN - 1) / TC_BN);
    slabs[i] = 0;
    for (int k = 0; k < g.seg_count; ++k) slabs[i] += (plan.segs[g.seg_begin + k].len + TC_BK - 1) / TC_BK;
  }
  const std::vector<int> ks = balance_split_factors(tiles, slabs);
  for (int i = 0; i < ng; ++i) {
    if (ks[i] < 2) continue;
    Group& g = plan.groups[i];
    float* p = arena->floats((size_t)ks[i] * g.M * g.N);   // ksplit planes: the reduce pass stays a valid fallback
    if (!p) continue;
    g.ksplit = ks[i];
    g.partial = p;
  }
}

// Copy the groups `idx` of `plan` (with their segments) into a new plan.
inline GemmPlan sub_plan(const GemmPlan& plan, const std::vector<int>& idx) {
  GemmPlan out;
  out.a_kmaj = plan.a_kmaj;
  out.b_kmaj = plan.b_kmaj;
  out.load_flags = plan.load_flags;
  out.label = plan.label;
  for (int i : idx) {
    Group g = plan.groups[i];
    const int b = g.seg_begin;
    g.seg_begin = (int)out.segs.size();
    for (int k = 0; k < g.seg_count; ++k) out.segs.push_back(plan.segs[b + k]);
    out.groups.push_back(g);
  }
  return out;
}

// Launch `plan` (all groups eligible) on the tcgen05 engine.
inline int launch_tc(const GemmPlan& plan_in, cudaStream_t stream, bool fixup = false, bool x3 = false) {
  // longest-K groups first (see the tile remap in the kernel): LPT-style balance of the tensor pipe
  std::vector<int> order(plan_in.groups.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    return plan_in.k_total(plan_in.groups[a]) / plan_in.groups[a].ksplit >
           plan_in.k_total(plan_in.groups[b]) / plan_in.groups[b].ksplit;
  });
  const GemmPlan plan = sub_plan(plan_in, order);
  // rank-3 maps for MN-major operands whose MN extent is a multiple of 32 in every group of the plan
  static const bool allow3d = []() {
    const char* e = getenv("TA3N_TMA_3D");
    return !(e && e[0] == '0');
  }();
  bool a3d = !plan.a_kmaj && allow3d, b3d = !plan.b_kmaj && allow3d;
  for (const Group& g : plan.groups) {
    if (g.M % 32 != 0) a3d = false;
    if (g.N % 32 != 0) b3d = false;
  }
  size_t gi = 0;
  while (gi < plan.groups.size()) {
    GemmTable tab;
    TcMaps maps;
    TcSegMaps sm;
    memset(&tab, 0, sizeof(int) * 4);
    memset(&sm, 0, sizeof(sm));
    std::map<MapKey, int> local;
    int ng = 0, ns = 0, tiles = 0, nmaps = 0, fix_tiles = 0;
    bool any_split = false;
    while (gi < plan.groups.size() && ng < kMaxGroups) {
      const Group& src = plan.groups[gi];
      if (ns + src.seg_count > kMaxSegs) break;
      // tensor maps of this group's segments (deduplicated within the launch)
      std::vector<std::pair<MapKey, MapKey>> keys;
      int fresh = 0;
      std::map<MapKey, int> trial = local;
      for (int i = 0; i < src.seg_count; ++i) {
        const Seg& s = plan.segs[src.seg_begin + i];
        const int raw = x3 ? 1 : 0;
        MapKey ka = plan.a_kmaj ? MapKey{s.A, s.len, src.M, s.lda, TC_BK, TC_BM, 0, 0, raw}
                                : MapKey{s.A, src.M, s.len, s.lda, 32, TC_BK, 1, a3d ? 1 : 0, raw};
        MapKey kb = plan.b_kmaj ? MapKey{s.B, s.len, src.N, s.ldb, TC_BK, TC_BN, 0, 0, raw}
                                : MapKey{s.B, src.N, s.len, s.ldb, 32, TC_BK, 1, b3d ? 1 : 0, raw};
        for (const MapKey& k : {ka, kb})
          if (!trial.count(k)) trial[k] = nmaps + fresh++;
        keys.push_back({ka, kb});
      }
      if (nmaps + fresh > kMaxMaps) {
        if (ng == 0) return fail(TA3N_ERR_UNSUPPORTED, "tcgen05 GEMM: one group needs %d tensor maps", fresh);
        break;
      }
      for (auto& kv : trial)
        if (!local.count(kv.first)) {
          TA3N_TRY(encode_map(kv.first, &maps.m[kv.second]));
          local[kv.first] = kv.second;
        }
      nmaps += fresh;
      Group g = src;
      for (int i = 0; i < src.seg_count; ++i) {
        tab.s[ns + i] = plan.segs[src.seg_begin + i];
        sm.a[ns + i] = (unsigned char)local[keys[i].first];
        sm.b[ns + i] = (unsigned char)local[keys[i].second];
      }
      g.seg_begin = ns;
      ns += src.seg_count;
      g.tiles_m = (g.M + TC_BM - 1) / TC_BM;
      g.tiles_n = (g.N + TC_BN - 1) / TC_BN;
      g.tile_begin = tiles;
      tiles += g.tiles_m * g.tiles_n * g.ksplit;
      g.fix_slot = -1;
      if (g.ksplit > 1 && fixup) {
        g.fix_slot = fix_tiles;                 // relative; rebased below once the ring slots are known
        fix_tiles += g.tiles_m * g.tiles_n;
      }
      tab.g[ng++] = g;
      ++gi;
    }
    // arrival counters for the fix-up groups; without them (first use inside a stream capture) the groups fall
    // back to the separate reduce pass
    int* fix_flags = nullptr;
    if (fix_tiles > 0) {
      int slot0 = 0;
      fix_flags = fix_flags_take(fix_tiles, stream, &slot0);
      for (int i = 0; i < ng; ++i)
        if (tab.g[i].fix_slot >= 0) tab.g[i].fix_slot = fix_flags ? tab.g[i].fix_slot + slot0 : -1;
    }
    for (int i = 0; i < ng; ++i) any_split |= tab.g[i].ksplit > 1 && tab.g[i].fix_slot < 0;
    tab.n_groups = ng;
    tab.total_tiles = tiles;
    tab.pad_ = (a3d ? 1 : 0) | (b3d ? 2 : 0);
    if (tiles > 0) {
      if (plan.a_kmaj && plan.b_kmaj)
        TA3N_TRY((tc_launch_one<true, true>(tab, maps, sm, stream, plan.label, fix_flags, x3)));
      else if (plan.a_kmaj && !plan.b_kmaj)
        TA3N_TRY((tc_launch_one<true, false>(tab, maps, sm, stream, plan.label, fix_flags, x3)));
      else if (!plan.a_kmaj && !plan.b_kmaj)
        TA3N_TRY((tc_launch_one<false, false>(tab, maps, sm, stream, plan.label, fix_flags, x3)));
      else
        TA3N_TRY((tc_launch_one<false, true>(tab, maps, sm, stream, plan.label, fix_flags, x3)));
      if (any_split) {
        dim3 grid(splitk_reduce_blocks(tab), ng);
        pre_launch("splitk_reduce", stream);
        launch_kernel(splitk_reduce_kernel, grid, 256, 0, stream, tab);
        TA3N_TRY(after_launch());
      }
    }
  }
  return TA3N_OK;
}

// Run a plan.  With the tf32 engine selected, every group the tensor-core kernel can take runs
// there; the rest (tiny heads, unaligned operands, ReLU-on-load) runs on the fp32 SIMT engine --
// still CUDA, never the CPU.
inline int run_gemm(GemmPlan& plan, cudaStream_t stream, Arena* splitk_arena = nullptr) {
  if (plan.groups.empty()) return TA3N_OK;
  for (auto& g : plan.groups)
    if (g.M <= 0 || g.N <= 0 || g.seg_count <= 0)
      return fail(TA3N_ERR_INVALID, "seg_gemm: empty group M=%d N=%d segs=%d", g.M, g.N, g.seg_count);
  const int engine = gemm_engine().load();
  if (engine == TA3N_GEMM_TF32_TCGEN05 || engine == TA3N_GEMM_TF32X3_TCGEN05) {
    const bool x3 = engine == TA3N_GEMM_TF32X3_TCGEN05;
    std::vector<int> tc_idx, simt_idx;
    for (int i = 0; i < (int)plan.groups.size(); ++i) (tc_group_ok(plan, plan.groups[i]) ? tc_idx : simt_idx).push_back(i);
    if (!tc_idx.empty()) {
      GemmPlan tc = sub_plan(plan, tc_idx);
      plan_splitk(tc, splitk_arena, TC_BM, TC_BN, TC_BK, 4);
      bool fixup = false;
      if (!x3 && fixup_enabled() && splitk_arena) {
        plan_balance_splitk(tc, splitk_arena);      // no-op when plan_splitk already split something
        for (const Group& g : tc.groups) fixup |= g.ksplit > 1;
      }
      TA3N_TRY(launch_tc(tc, stream, fixup, x3));
    }
    if (!simt_idx.empty()) {
      GemmPlan rest = sub_plan(plan, simt_idx);
      plan_splitk(rest, splitk_arena, SG_BM, SG_BN, SG_BK, 8);
      TA3N_TRY(launch_simt(rest, stream));
    }
    return TA3N_OK;
  }
  plan_splitk(plan, splitk_arena, SG_BM, SG_BN, SG_BK, 8);
  return launch_simt(plan, stream);
}

}  // namespace ta3n
