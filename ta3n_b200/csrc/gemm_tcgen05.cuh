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
#include <cmath>
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
constexpr int TC_TMEM_COLS = 128;      // per-launch kernel: one accumulator; the step kernel allocates two
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

// ---- one output tile, by warp role ---------------------------------------------------------------------
// Shared by the per-launch kernel below (one tile per CTA, all roles in lockstep) and by the persistent step kernel
// (step_kernel.cuh), where the three roles run their own loops over the task queue: the producer and the MMA issuer
// of tile t+1 work while the epilogue warps still drain tile t (two TMEM accumulator buffers).
// The operand ring barriers are initialised once per CTA; each role keeps its own running slab count.
struct TcShared {
  uint64_t full_bar[6];
  uint64_t empty_bar[6];
  uint64_t tmem_full_bar[2];
  uint64_t tmem_empty_bar[2];
  uint32_t tmem_slot;
};

template <int TC_STAGES>
__device__ __forceinline__ void tc_pipe_init(TcShared* sh, int epi_warps) {
  for (int s = 0; s < TC_STAGES; ++s) {
    mbar_init(&sh->full_bar[s], 1);
    mbar_init(&sh->empty_bar[s], 1);
  }
  for (int a = 0; a < 2; ++a) {
    mbar_init(&sh->tmem_full_bar[a], 1);
    mbar_init(&sh->tmem_empty_bar[a], epi_warps);
  }
  fence_barrier_init();
}

// What a tile of a split-K group does with its accumulator:
//   TILE_FINAL   : fused epilogue -> C                       (ksplit == 1)
//   TILE_PARTIAL : raw accumulator -> partial[split]         (separate reduce pass, or an owner tile below)
//   TILE_OWNER   : adds partial[0 .. ksplit-2] in a fixed order, then fused epilogue -> C.  The caller guarantees
//                  those partials are complete and visible (step kernel: task dependency).
//   TILE_SPLIT   : (step kernel) raw accumulator -> partial[split]; the LAST split of the tile to arrive (arrival counter)
//                  then runs a TILE_REDUCE pass: partial[0 .. ksplit-1] summed in split order + fused epilogue -> C.
//                  No split ever waits for another one, and the result does not depend on which one came last.
enum : int { TILE_FINAL = 0, TILE_PARTIAL = 1, TILE_OWNER = 2, TILE_SPLIT = 3, TILE_REDUCE = 4 };

// TMA producer (ONE thread): the n_iter K slabs [c_begin, c_begin + n_iter) of tile (m0, n0) into the ring.
// `slabs` = slabs this CTA has pushed so far (advanced by the caller).
template <int TC_STAGES>
__device__ __forceinline__ void tc_produce(const TileCtx& ctx, const CUtensorMap* __restrict__ maps, const bool a_kmaj,
                                           const bool b_kmaj, const int pad_flags, const int m0, const int n0,
                                           const int c_begin, const int n_iter, uint8_t* smem, TcShared* sh,
                                           const uint32_t slabs) {
  const Group& g = ctx.g;
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
  for (int it = 0; it < n_iter; ++it) {
    const uint32_t gl = slabs + (uint32_t)it;
    const int stage = (int)(gl % TC_STAGES);
    const uint32_t phase = (gl / TC_STAGES) & 1u;
    mbar_wait(&sh->empty_bar[stage], phase ^ 1u);
    mbar_expect_tx(&sh->full_bar[stage], TC_STAGE_BYTES);
    uint8_t* sA = smem + stage * TC_STAGE_BYTES;
    uint8_t* sB = sA + TC_A_BYTES;
    const CUtensorMap* ma = &maps[ctx.seg[seg].amap];
    const CUtensorMap* mb = &maps[ctx.seg[seg].bmap];
    // MN-major tiles are four [32 k-rows][32 floats] slabs, one per group of 32 m (or n).  When the operand's
    // MN extent is a multiple of 32 a rank-3 tensor map {32 floats, k rows, groups of 32} fetches all four with
    // ONE instruction (the lone producer thread is issue-bound: ~50 cycles per TMA, 8 per chunk otherwise).
    if (a_kmaj) {
      tma_load_2d(sA, ma, &sh->full_bar[stage], k0, m0);
    } else if (pad_flags & 1) {
      tma_load_3d(sA, ma, &sh->full_bar[stage], 0, k0, m0 >> 5);
    } else {
#pragma unroll
      for (int q = 0; q < TC_BM / 32; ++q) tma_load_2d(sA + q * 4096, ma, &sh->full_bar[stage], m0 + 32 * q, k0);
    }
    if (b_kmaj) {
      tma_load_2d(sB, mb, &sh->full_bar[stage], k0, n0);
    } else if (pad_flags & 2) {
      tma_load_3d(sB, mb, &sh->full_bar[stage], 0, k0, n0 >> 5);
    } else {
#pragma unroll
      for (int q = 0; q < TC_BN / 32; ++q) tma_load_2d(sB + q * 4096, mb, &sh->full_bar[stage], n0 + 32 * q, k0);
    }
    k0 += TC_BK;
    if (k0 >= ctx.seg[seg].len) {
      ++seg;
      k0 = 0;
    }
  }
}

// MMA issuer (ONE thread): accumulates the n_iter slabs into TMEM buffer `acc` (columns acc*128 ..).
template <int TC_STAGES>
__device__ __forceinline__ void tc_mma(const bool a_kmaj, const bool b_kmaj, const int n_iter, uint8_t* smem, TcShared* sh,
                                       const uint32_t tmem_base, const int acc, const uint32_t slabs) {
  const uint32_t idesc = umma_idesc_tf32(a_kmaj, b_kmaj, TC_BM, TC_BN);
  const uint32_t tmem_d = tmem_base + (uint32_t)(acc * TC_BN);
  for (int it = 0; it < n_iter; ++it) {
    const uint32_t gl = slabs + (uint32_t)it;
    const int stage = (int)(gl % TC_STAGES);
    const uint32_t phase = (gl / TC_STAGES) & 1u;
    mbar_wait(&sh->full_bar[stage], phase);
    tc_fence_after();
    const uint32_t a_base = smem_u32(smem + stage * TC_STAGE_BYTES);
    const uint32_t b_base = a_base + TC_A_BYTES;
#pragma unroll
    for (int ks = 0; ks < TC_BK / 8; ++ks) {
      const uint64_t adesc = a_kmaj ? umma_desc_kmajor(a_base, ks) : umma_desc_mnmajor(a_base, ks);
      const uint64_t bdesc = b_kmaj ? umma_desc_kmajor(b_base, ks) : umma_desc_mnmajor(b_base, ks);
      umma_tf32(tmem_d, adesc, bdesc, idesc, (it > 0 || ks > 0) ? 1u : 0u);
    }
    umma_commit(&sh->empty_bar[stage]);   // releases the smem slot once these MMAs have read it
  }
  umma_commit(&sh->tmem_full_bar[acc]);   // accumulator complete
}

// Epilogue (kEpiWarps warps; `ew` = this warp's index among them, 0-based; warp id % 4 must equal (ew + 2) % 4, i.e.
// the epilogue warps are warps 2 .. 2 + kEpiWarps - 1 of the CTA).  Lane i of a warp owns accumulator row 32*lq + i;
// tcgen05.ld hands it 32 consecutive columns, which it finishes (fused epilogue) and writes as 8 x 16 B stores:
// every 128 B line of C is written whole.  The caller has waited for tmem_full.
template <int kEpiWarps>
__device__ __forceinline__ void tc_epilogue(const TileCtx& ctx, const int m0, const int n0, const int split,
                                            const int n_iter, const int mode, const uint32_t tmem_base, const int acc,
                                            const int ew) {
  const int lane = threadIdx.x & 31;
  const int lq = (ew + 2) & 3;          // TMEM lane quarter this warp may access
  const Group e = ctx.g;                // register copy: no reloads behind the global stores
  const int m = m0 + lq * 32 + lane;
  const bool split_out = mode == TILE_PARTIAL;
  float* const obase = split_out ? e.partial + (size_t)split * e.M * e.N : e.C;
  const int ldo = split_out ? e.N : e.ldc;
  constexpr int kColChunks = (TC_BN / 32) * 4 / kEpiWarps;      // column chunks of 32 per warp: 4 or 2
  const int c0 = (ew / 4) * kColChunks;
#pragma unroll 1
  for (int c = c0; c < c0 + kColChunks; ++c) {
    float v[32];
    if (n_iter > 0) {
      tmem_ld_32x32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(acc * TC_BN + c * 32), v);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
    const int nb = n0 + c * 32;
    if (m < e.M && nb < e.N) {
      float* orow = obase + (size_t)m * ldo + nb;
      const int nvalid = min(32, e.N - nb);
      if (mode == TILE_OWNER) {      // raw partial sums of the other splits, two in flight, added in split order
        float pr[2][32];
        load_row32(e.partial + (size_t)m * e.N + nb, nvalid, pr[0]);
        if (e.ksplit > 2) load_row32(e.partial + (size_t)e.M * e.N + (size_t)m * e.N + nb, nvalid, pr[1]);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += pr[0][j];
        if (e.ksplit > 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += pr[1][j];
        }
        for (int sp = 2; sp < e.ksplit - 1; ++sp) {
          load_row32(e.partial + (size_t)sp * e.M * e.N + (size_t)m * e.N + nb, nvalid, pr[0]);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += pr[0][j];
        }
      }
      if (!split_out) {
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
  tc_fence_before();                    // TMEM reads ordered before whatever lets the next MMA reuse the buffer
}

// ---- coalesced epilogue (step kernel) -------------------------------------------------------------------
// The epilogue above gives every lane one accumulator ROW: each of its memory instructions touches 32 different
// 128 B lines, i.e. 32 LSU cycles per instruction -- 2 us per plain tile and 20-45 us for the tiles whose epilogue
// reads several auxiliary operands (measured: step trace, relation-discriminator data gradient).  Here a warp stages
// its 32 x 32 accumulator chunk in shared memory ([32][36] floats, conflict-free for the 16 B accesses of both
// passes) and re-reads it with lanes along the COLUMNS: lane = (row quarter rq, column quad cq), every global access
// is a 16 B piece of a 128 B row segment, 4 rows per instruction.  Auxiliary operands of four rows are requested
// together before the first is used.
// ONE body with run-time flags (uniform branches), not one copy per flag set: the nine specialised copies of the
// first version made the step kernel 42 k instructions (680 KB), far beyond the instruction cache.  The step planner
// guarantees what keeps it small: N % 4 == 0, every leading dimension % 4 == 0, every pointer 16-byte aligned (no
// scalar tails), ksplit <= 4, and no auxiliary operand on a split group (the three aux registers sets are shared
// between {partials} and {add, gate, accumulate}).
constexpr int TC_STAGE_LD = 36;
constexpr int TC_EPI_STAGE_FLOATS = 32 * TC_STAGE_LD;      // per epilogue warp

__device__ __forceinline__ void sts4(uint32_t saddr, const float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void add4(float4& a, const float4& b) {
  a.x += b.x;
  a.y += b.y;
  a.z += b.z;
  a.w += b.w;
}

// CLS selects how much of the body exists (three copies instead of one per flag set):
//   0 = plain: alpha * accumulator (+ split-K partials)          weight gradients, partial tiles, frame dgrad
//   1 = forward: + bias, ReLU, dropout                            every forward layer
//   2 = everything (run-time flags)                               the data-gradient tiles with auxiliary operands
// The per-element work of the plain class is one multiply: with run-time flag tests inside the row loops it was
// 1300 instructions per warp and tile (5 us at 2 warps per scheduler); the lean copies are ~250.
enum : int { EPI_CLS_PLAIN = 0, EPI_CLS_FORWARD = 1, EPI_CLS_ALL = 2 };
__device__ __forceinline__ int epi_class(const int mode, const int flags) {
  if (mode == TILE_PARTIAL || mode == TILE_SPLIT || flags == 0) return EPI_CLS_PLAIN;
  if ((flags & ~(EPI_BIAS | EPI_RELU | EPI_DROP_MASK | EPI_DROP_RNG)) == 0) return EPI_CLS_FORWARD;
  return EPI_CLS_ALL;
}

template <int kEpiWarps, int CLS>
__device__ __forceinline__ void tc_epilogue_cls(const TileCtx& ctx, const int m0, const int n0, const int split,
                                                const int n_iter, const int mode, const uint32_t tmem_base,
                                                const int acc, const int ew, const uint32_t stage) {
  const int lane = threadIdx.x & 31;
  const int lq = (ew + 2) & 3;          // TMEM lane quarter this warp may access
  const Group& e = ctx.g;               // shared memory (the task slot)
  constexpr int kMask = CLS == EPI_CLS_PLAIN ? 0
                        : CLS == EPI_CLS_FORWARD ? (EPI_BIAS | EPI_RELU | EPI_DROP_MASK | EPI_DROP_RNG)
                                                 : ~0;
  const bool split_out = mode == TILE_PARTIAL || mode == TILE_SPLIT;
  const int f = (split_out ? 0 : e.flags) & kMask;
  const int M = e.M, N = e.N;
  const size_t plane = (size_t)M * N;
  float* const obase = split_out ? e.partial + (size_t)split * plane : e.C;
  const int ldo = split_out ? N : e.ldc;
  constexpr int kColChunks = (TC_BN / 32) * 4 / kEpiWarps;
  const int c0 = (ew / 4) * kColChunks;
  const int rq = lane >> 3, cq = lane & 7;
  // planes of raw partial sums folded in (no aux operands on split groups: the register sets are shared)
  const int n_part = CLS == EPI_CLS_ALL ? 0 : (mode == TILE_OWNER ? e.ksplit - 1 : (mode == TILE_REDUCE ? e.ksplit : 0));
  const float alpha = split_out ? 1.0f : (e.alpha_dev ? e.alpha * __ldg(e.alpha_dev) : e.alpha);
  const uint64_t step = (f & EPI_DROP_RNG) ? (e.step_dev ? *e.step_dev : 0ull) : 0ull;
  const bool drop_early = (f & (EPI_DROP_MASK | EPI_DROP_RNG)) && !(f & EPI_DROP_LATE);
  const bool drop_late = (f & (EPI_DROP_MASK | EPI_DROP_RNG)) && (f & EPI_DROP_LATE);
#pragma unroll 1
  for (int c = c0; c < c0 + kColChunks; ++c) {
    {  // accumulator chunk -> registers (lane = row) -> shared memory
      float v[32];
      if (n_iter > 0) {
        tmem_ld_32x32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(acc * TC_BN + c * 32), v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      __syncwarp();                     // the previous chunk's readers are done with the staging tile
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        sts4(stage + (uint32_t)(lane * TC_STAGE_LD + j) * 4u, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
      __syncwarp();
    }
    const int n = n0 + c * 32 + cq * 4;
    if (n >= N) continue;               // N % 4 == 0: a column quad is inside or outside as a whole
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f & EPI_BIAS) bias4 = ldcg4(e.bias + n);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {                 // 2 batches of 4 x (4 rows per instruction) = 32 rows
      int rows[4];
      bool ok[4];
      float4 q[4], x0[4], x1[4], x2[4], x3[4];
      float rs[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = (half * 4 + u) * 4 + rq;            // row of the chunk
        rows[u] = m0 + lq * 32 + r;
        ok[u] = rows[u] < M;
        q[u] = lds4(stage + (uint32_t)(r * TC_STAGE_LD + cq * 4) * 4u);
      }
      // ---- every load of the batch first ----
      // (every element of the auxiliary arrays is written, under a select rather than a branch: left partly
      //  uninitialised behind `continue`s the compiler kept them in LOCAL memory and stored each load result to the
      //  stack as it arrived -- one load in flight at a time; measured with ncu, profiles/README.md)
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t m = (size_t)(ok[u] ? rows[u] : m0);        // a safe row for the masked-off lanes
        x0[u] = x1[u] = x2[u] = zero4;
        rs[u] = 1.0f;
        if (n_part > 0) {
          x0[u] = ldcg4(e.partial + m * N + n);
          if (n_part > 1) x1[u] = ldcg4(e.partial + plane + m * N + n);
          if (n_part > 2) x2[u] = ldcg4(e.partial + 2 * plane + m * N + n);
          x3[u] = n_part > 3 ? ldcg4(e.partial + 3 * plane + m * N + n) : zero4;
        } else if (CLS == EPI_CLS_ALL) {
          if (f & EPI_ADDROW) {
            if (e.rowscale) rs[u] = __ldcg(e.rowscale + m * e.rs_stride) + e.rs_bias;
            x0[u] = ldcg4(e.add + m * e.ldadd + n);
          }
          if (f & (EPI_GATE | EPI_DPRE)) x1[u] = ldcg4(e.gate + m * e.ldgate + n);
          if (f & EPI_ACCUM) x2[u] = ldcg4(e.C + m * e.ldc + n);
          if (f & EPI_MULTI) {          // the gates of every dZ plane ride with the first round of loads
            x1[u] = ldcg4(e.multi_gate[0] + m * e.ldmulti + n);           // (a MULTI group has no gate / accumulate)
            if (e.n_multi > 1) x2[u] = ldcg4(e.multi_gate[1] + m * e.ldmulti + n);
            x3[u] = e.n_multi > 2 ? ldcg4(e.multi_gate[2] + m * e.ldmulti + n) : zero4;
          }
        }
      }
      // ---- arithmetic + store ----
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t m = (size_t)(ok[u] ? rows[u] : m0);
        float4 v = q[u];
        if (n_part > 0) {
          add4(v, x0[u]);
          if (n_part > 1) add4(v, x1[u]);
          if (n_part > 2) add4(v, x2[u]);
          if (n_part > 3) add4(v, x3[u]);
        }
        if (!split_out) {
          float ev[4] = {v.x * alpha, v.y * alpha, v.z * alpha, v.w * alpha};
          if (f & EPI_BIAS) {
            ev[0] += bias4.x;
            ev[1] += bias4.y;
            ev[2] += bias4.z;
            ev[3] += bias4.w;
          }
          if (f & EPI_RELU) {
#pragma unroll
            for (int i = 0; i < 4; ++i) ev[i] = fmaxf(ev[i], 0.f);
          }
          if (drop_early || drop_late) {
            float df[4];
            if (f & EPI_DROP_MASK) {
              const uchar4 k = *reinterpret_cast<const uchar4*>(e.keep + m * e.ldkeep + n);
              df[0] = k.x ? e.drop_scale : 0.f;
              df[1] = k.y ? e.drop_scale : 0.f;
              df[2] = k.z ? e.drop_scale : 0.f;
              df[3] = k.w ? e.drop_scale : 0.f;
            } else {
              const uint64_t base = e.rng_offset + m * (uint64_t)N + (uint64_t)n;
              if ((base & 3ull) == 0) {                     // the quad shares one hash (always, on the path's shapes)
                const uint64_t hsh = rng_hash4(e.seed, step, base >> 2);
                const uint32_t thr = rng_threshold(e.drop_p);
#pragma unroll
                for (int i = 0; i < 4; ++i) df[i] = rng_keep_bits(hsh, i, thr) ? e.drop_scale : 0.f;
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) df[i] = rng_keep(e.seed, step, base + i, e.drop_p) ? e.drop_scale : 0.f;
              }
            }
            if (drop_late && (f & EPI_ADDROW)) {
              ev[0] = fmaf(rs[u], x0[u].x, ev[0]);
              ev[1] = fmaf(rs[u], x0[u].y, ev[1]);
              ev[2] = fmaf(rs[u], x0[u].z, ev[2]);
              ev[3] = fmaf(rs[u], x0[u].w, ev[3]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) ev[i] = df[i] != 0.f ? ev[i] * df[i] : 0.f;
          }
          if ((f & EPI_ADDROW) && !drop_late) {
            ev[0] = fmaf(rs[u], x0[u].x, ev[0]);
            ev[1] = fmaf(rs[u], x0[u].y, ev[1]);
            ev[2] = fmaf(rs[u], x0[u].z, ev[2]);
            ev[3] = fmaf(rs[u], x0[u].w, ev[3]);
          }
          if (f & EPI_GATE) {
            ev[0] = x1[u].x > 0.f ? ev[0] : 0.f;
            ev[1] = x1[u].y > 0.f ? ev[1] : 0.f;
            ev[2] = x1[u].z > 0.f ? ev[2] : 0.f;
            ev[3] = x1[u].w > 0.f ? ev[3] : 0.f;
          }
          if (f & EPI_ACCUM) {
            ev[0] += x2[u].x;
            ev[1] += x2[u].y;
            ev[2] += x2[u].z;
            ev[3] += x2[u].w;
          }
          if (f & EPI_DPRE) {
            ev[0] = x1[u].x > 0.f ? ev[0] * e.drop_scale : 0.f;
            ev[1] = x1[u].y > 0.f ? ev[1] * e.drop_scale : 0.f;
            ev[2] = x1[u].z > 0.f ? ev[2] * e.drop_scale : 0.f;
            ev[3] = x1[u].w > 0.f ? ev[3] * e.drop_scale : 0.f;
          }
          v = make_float4(ev[0], ev[1], ev[2], ev[3]);
        }
        q[u] = v;
        if (ok[u]) *reinterpret_cast<float4*>(obase + m * ldo + n) = v;
      }
      if (CLS == EPI_CLS_ALL && (f & EPI_MULTI)) {      // dZ planes (gates already in x1 .. x3)
        const int n_multi = e.n_multi;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (ok[u]) {
            const size_t off = (size_t)rows[u] * e.ldmulti + n;
            *reinterpret_cast<float4*>(e.multi_out[0] + off) =
                make_float4(x1[u].x > 0.f ? q[u].x : 0.f, x1[u].y > 0.f ? q[u].y : 0.f, x1[u].z > 0.f ? q[u].z : 0.f,
                            x1[u].w > 0.f ? q[u].w : 0.f);
            if (n_multi > 1)
              *reinterpret_cast<float4*>(e.multi_out[1] + off) =
                  make_float4(x2[u].x > 0.f ? q[u].x : 0.f, x2[u].y > 0.f ? q[u].y : 0.f, x2[u].z > 0.f ? q[u].z : 0.f,
                              x2[u].w > 0.f ? q[u].w : 0.f);
            if (n_multi > 2)
              *reinterpret_cast<float4*>(e.multi_out[2] + off) =
                  make_float4(x3[u].x > 0.f ? q[u].x : 0.f, x3[u].y > 0.f ? q[u].y : 0.f, x3[u].z > 0.f ? q[u].z : 0.f,
                              x3[u].w > 0.f ? q[u].w : 0.f);
          }
      }
    }
  }
  tc_fence_before();
}

// Can the step kernel's epilogue take this group?  (see tc_epilogue_coalesced)
inline bool tc_step_group_ok(const Group& g) {
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (g.N % 4 != 0 || g.ldc % 4 != 0 || !a16(g.C) || g.ksplit > 4) return false;
  if (g.ksplit > 1 && (g.flags & (EPI_ADDROW | EPI_GATE | EPI_ACCUM | EPI_DPRE | EPI_MULTI))) return false;
  if ((g.flags & EPI_BIAS) && !a16(g.bias)) return false;
  if ((g.flags & EPI_DROP_MASK) && (g.ldkeep % 4 != 0 || (reinterpret_cast<uintptr_t>(g.keep) & 3u) != 0)) return false;
  if ((g.flags & EPI_ADDROW) && (g.ldadd % 4 != 0 || !a16(g.add))) return false;
  if ((g.flags & (EPI_GATE | EPI_DPRE)) && (g.ldgate % 4 != 0 || !a16(g.gate))) return false;
  if (g.flags & EPI_MULTI) {
    if (g.ldmulti % 4 != 0 || g.n_multi < 1 || g.n_multi > 3) return false;
    if (g.flags & (EPI_GATE | EPI_DPRE | EPI_ACCUM)) return false;      // their registers carry the plane gates
    for (int p = 0; p < g.n_multi; ++p)
      if (!a16(g.multi_gate[p]) || !a16(g.multi_out[p])) return false;
  }
  return true;
}

// chunk range [c_begin, c_begin + n_iter) of split `split` of the group staged in ctx
__device__ __forceinline__ void tc_chunk_range(const TileCtx& ctx, int split, int* c_begin, int* n_iter) {
  const Group& g = ctx.g;
  int total_chunks = 0;
  for (int s = 0; s < g.seg_count; ++s) total_chunks += (ctx.seg[s].len + TC_BK - 1) / TC_BK;
  const int cps = (total_chunks + g.ksplit - 1) / g.ksplit;
  *c_begin = split * cps;
  *n_iter = max(0, min(total_chunks, *c_begin + cps) - *c_begin);
}

// ---- the per-launch kernel: one tile per CTA ---------------------------------------------------------
template <bool A_KMAJ, bool B_KMAJ, int TC_STAGES>
__global__ void __launch_bounds__(tc_threads(TC_STAGES), TC_STAGES > 3 ? 1 : 2)
seg_gemm_tc_kernel(const __grid_constant__ GemmTable tab, const __grid_constant__ TcMaps maps,
                   const __grid_constant__ TcSegMaps segmaps, const int first_wave) {
  extern __shared__ uint8_t tc_smem_raw[];
  __shared__ __align__(8) TcShared sh;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- tile decode (same scheme as the SIMT engine, 128x128 tiles); group + segments staged in smem ----
  __shared__ TileCtx ctx;
  // Groups arrive sorted by K (longest first).  CTAs beyond the first wave (one per SM) take tiles from the END
  // of the list, so the SM that received the longest tile gets the shortest one as its second resident CTA.
  int tile = blockIdx.x;
  if (tile >= first_wave) tile = tab.total_tiles - 1 - (tile - first_wave);
  load_tile_ctx(tab, tile, &ctx, segmaps.a, segmaps.b);
  const Group& g = ctx.g;
  int local = tile - g.tile_begin;
  const int per_split = g.tiles_m * g.tiles_n;
  const int split = local / per_split;
  local -= split * per_split;
  const int m0 = (local / g.tiles_n) * TC_BM;
  const int n0 = (local % g.tiles_n) * TC_BN;
  int c_begin, n_iter;
  tc_chunk_range(ctx, split, &c_begin, &n_iter);

  // ---- one-time setup ----
  constexpr int kEpiWarps = tc_threads(TC_STAGES) / 32 - 2;
  if (warp == 0 && lane == 0) tc_pipe_init<TC_STAGES>(&sh, kEpiWarps);
  if (warp == 1) tmem_alloc(&sh.tmem_slot, TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_slot;
  // Everything above touched only kernel parameters, shared memory and TMEM: it overlaps the previous
  // kernel of the stream.  From here on operands produced by that kernel are read.
  pdl_wait();

  const int mode = g.ksplit > 1 ? TILE_PARTIAL : TILE_FINAL;
  if (warp == 0) {
    if (lane == 0 && n_iter > 0) tc_produce<TC_STAGES>(ctx, maps.m, A_KMAJ, B_KMAJ, tab.pad_, m0, n0, c_begin, n_iter, smem, &sh, 0u);
  } else if (warp == 1) {
    if (lane == 0 && n_iter > 0) tc_mma<TC_STAGES>(A_KMAJ, B_KMAJ, n_iter, smem, &sh, tmem_base, 0, 0u);
  } else {
    if (n_iter > 0) {
      mbar_wait(&sh.tmem_full_bar[0], 0u);
      tc_fence_after();
    }
    tc_epilogue<kEpiWarps>(ctx, m0, n0, split, n_iter, mode, tmem_base, 0, warp - 2);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

// ---- the precise forward kernel ("tf32x3"): fp32-grade products from the tensor cores -------------------------
// kind::tf32 keeps 10 mantissa bits of each operand: every forward GEMM is accurate to ~3e-4, and ~1.3e-4 of the
// ReLU units behind them (pre-activation within that error of zero) come out with the wrong on/off state.  Each flip
// moves a gradient entry by O(1), which is where the 1-2 % gradient deviation of the plain tf32 engine comes from
// (tools/parity_report.py; with the pattern pinned the same gradients agree to 4e-4).  The forward layers therefore
// run here with the operands split in two tf32 pieces (a = hi + lo, hi = the 10 mantissa bits the tensor core reads
// from the raw fp32 word, lo = a - hi, exact):     a b ~= hi_a hi_b + lo_a hi_b + hi_a lo_b      (three MMAs per K step)
// Two more things are needed for fp32 grade (measured, tools/x3_probe.py, profiles/README.md):
//   * K is accumulated in CHUNKS of 256 inside the tensor core and the chunks are added in fp32 registers: over the
//     full K = 2048 the tensor core's own accumulation limits the result to 5e-6 whatever the operands (three and four
//     products give the same error); chunked, 6e-7 -- the error of an fp32 FFMA loop -- and no sign flips;
//   * the tensor maps deliver the RAW fp32 words (no TMA rounding), so that hi + lo = a exactly.
// The eight warps that wait for the accumulator in the plain kernel do the work: they compute the lo tiles of each
// stage (shared memory -> shared memory, layout-agnostic: lo has the swizzle of its source), and drain a finished chunk
// from TMEM (two accumulator buffers: the MMAs of chunk c+1 run while chunk c is added to the 64 running sums a thread
// keeps), then run the usual fused epilogue on those registers.
constexpr int X3_STAGES = 3;
constexpr int X3_STAGE_BYTES = 2 * TC_STAGE_BYTES;            // [A | B | A_lo | B_lo]
constexpr int X3_CHUNK = 8;                                   // slabs (256 K columns) per tensor-core accumulation
constexpr int X3_THREADS = 320;                               // warp 0 TMA, warp 1 MMA, 8 worker warps
constexpr int X3_TMEM_COLS = 256;
constexpr int x3_smem_bytes() { return X3_STAGES * X3_STAGE_BYTES + 1024; }

struct X3Shared {
  uint64_t full_bar[X3_STAGES];      // TMA landed
  uint64_t split_bar[X3_STAGES];     // the 8 worker warps have written the lo tiles
  uint64_t empty_bar[X3_STAGES];     // the MMAs have read the stage
  uint64_t chunk_full[2];            // accumulator buffer complete
  uint64_t chunk_empty[2];           // ... drained by the 8 worker warps
  uint32_t tmem_slot;
};

template <bool A_KMAJ, bool B_KMAJ>
__global__ void __launch_bounds__(X3_THREADS, 1)
seg_gemm_tc_x3_kernel(const __grid_constant__ GemmTable tab, const __grid_constant__ TcMaps maps,
                      const __grid_constant__ TcSegMaps segmaps) {
  extern __shared__ uint8_t tc_smem_raw[];
  __shared__ __align__(8) X3Shared sh;
  __shared__ TileCtx ctx;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int tile = blockIdx.x;
  load_tile_ctx(tab, tile, &ctx, segmaps.a, segmaps.b);
  const Group& g = ctx.g;
  int local = tile - g.tile_begin;
  const int per_split = g.tiles_m * g.tiles_n;
  const int split = local / per_split;
  local -= split * per_split;
  const int m0 = (local / g.tiles_n) * TC_BM;
  const int n0 = (local % g.tiles_n) * TC_BN;
  int c_begin, n_iter;
  tc_chunk_range(ctx, split, &c_begin, &n_iter);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < X3_STAGES; ++s) {
      mbar_init(&sh.full_bar[s], 1);
      mbar_init(&sh.split_bar[s], 8);
      mbar_init(&sh.empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sh.chunk_full[b], 1);
      mbar_init(&sh.chunk_empty[b], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&sh.tmem_slot, X3_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_slot;
  pdl_wait();

  const int n_chunks = (n_iter + X3_CHUNK - 1) / X3_CHUNK;
  if (warp == 0) {
    // ---------------- TMA producer ----------------
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
      for (int it = 0; it < n_iter; ++it) {
        const int stage = it % X3_STAGES;
        const uint32_t phase = (uint32_t)(it / X3_STAGES) & 1u;
        mbar_wait(&sh.empty_bar[stage], phase ^ 1u);
        mbar_expect_tx(&sh.full_bar[stage], TC_STAGE_BYTES);
        uint8_t* sA = smem + stage * X3_STAGE_BYTES;
        uint8_t* sB = sA + TC_A_BYTES;
        const CUtensorMap* ma = &maps.m[ctx.seg[seg].amap];
        const CUtensorMap* mb = &maps.m[ctx.seg[seg].bmap];
        if (A_KMAJ) {
          tma_load_2d(sA, ma, &sh.full_bar[stage], k0, m0);
        } else if (tab.pad_ & 1) {
          tma_load_3d(sA, ma, &sh.full_bar[stage], 0, k0, m0 >> 5);
        } else {
#pragma unroll
          for (int q = 0; q < TC_BM / 32; ++q) tma_load_2d(sA + q * 4096, ma, &sh.full_bar[stage], m0 + 32 * q, k0);
        }
        if (B_KMAJ) {
          tma_load_2d(sB, mb, &sh.full_bar[stage], k0, n0);
        } else if (tab.pad_ & 2) {
          tma_load_3d(sB, mb, &sh.full_bar[stage], 0, k0, n0 >> 5);
        } else {
#pragma unroll
          for (int q = 0; q < TC_BN / 32; ++q) tma_load_2d(sB + q * 4096, mb, &sh.full_bar[stage], n0 + 32 * q, k0);
        }
        k0 += TC_BK;
        if (k0 >= ctx.seg[seg].len) {
          ++seg;
          k0 = 0;
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: three products per K step, a fresh accumulator per chunk ----------------
    if (lane == 0 && n_iter > 0) {
      const uint32_t idesc = umma_idesc_tf32(A_KMAJ, B_KMAJ, TC_BM, TC_BN);
      for (int it = 0; it < n_iter; ++it) {
        const int stage = it % X3_STAGES;
        const uint32_t phase = (uint32_t)(it / X3_STAGES) & 1u;
        const int chunk = it / X3_CHUNK, cbuf = chunk & 1;
        const bool chunk_first = it % X3_CHUNK == 0;
        if (chunk_first) {
          mbar_wait(&sh.chunk_empty[cbuf], ((uint32_t)(chunk >> 1) & 1u) ^ 1u);      // drained two chunks ago
          tc_fence_after();
        }
        mbar_wait(&sh.split_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + stage * X3_STAGE_BYTES);
        const uint32_t b_hi = a_hi + TC_A_BYTES;
        const uint32_t a_lo = a_hi + TC_STAGE_BYTES, b_lo = b_hi + TC_STAGE_BYTES;
        const uint32_t tmem_d = tmem_base + (uint32_t)(cbuf * TC_BN);
#pragma unroll
        for (int ks = 0; ks < TC_BK / 8; ++ks) {
          const uint64_t dah = A_KMAJ ? umma_desc_kmajor(a_hi, ks) : umma_desc_mnmajor(a_hi, ks);
          const uint64_t dbh = B_KMAJ ? umma_desc_kmajor(b_hi, ks) : umma_desc_mnmajor(b_hi, ks);
          const uint64_t dal = A_KMAJ ? umma_desc_kmajor(a_lo, ks) : umma_desc_mnmajor(a_lo, ks);
          const uint64_t dbl = B_KMAJ ? umma_desc_kmajor(b_lo, ks) : umma_desc_mnmajor(b_lo, ks);
          umma_tf32(tmem_d, dal, dbh, idesc, (chunk_first && ks == 0) ? 0u : 1u);      // small terms first
          umma_tf32(tmem_d, dah, dbl, idesc, 1u);
          umma_tf32(tmem_d, dah, dbh, idesc, 1u);
        }
        umma_commit(&sh.empty_bar[stage]);
        if (it % X3_CHUNK == X3_CHUNK - 1 || it == n_iter - 1) umma_commit(&sh.chunk_full[cbuf]);
      }
    }
  } else {
    // ---------------- worker warps: lo tiles, chunk accumulation, epilogue ----------------
    const int ew = warp - 2;
    const int wt = ew * 32 + lane;                      // 0 .. 255
    const int lq = (ew + 2) & 3;                        // TMEM lane quarter of this warp
    const int chalf = ew / 4;                           // which 64 columns of the tile
    float sum[2][32];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int j = 0; j < 32; ++j) sum[c][j] = 0.f;
    int drained = 0;
    auto drain = [&](int chunk) {
      const int cbuf = chunk & 1;
      mbar_wait(&sh.chunk_full[cbuf], (uint32_t)(chunk >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(cbuf * TC_BN + (chalf * 2 + c) * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) sum[c][j] += v[j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.chunk_empty[cbuf]);
    };
    for (int it = 0; it < n_iter; ++it) {
      const int stage = it % X3_STAGES;
      const uint32_t phase = (uint32_t)(it / X3_STAGES) & 1u;
      mbar_wait(&sh.full_bar[stage], phase);
      const uint32_t hi = smem_u32(smem + stage * X3_STAGE_BYTES);
#pragma unroll
      for (int i = 0; i < TC_STAGE_BYTES / 16 / 256; ++i) {       // 8 float4 per thread: A and B tiles alike
        const uint32_t off = (uint32_t)(wt + 256 * i) * 16u;
        const float4 a = lds4(hi + off);
        float4 l;
        l.x = a.x - __uint_as_float(__float_as_uint(a.x) & 0xFFFFE000u);
        l.y = a.y - __uint_as_float(__float_as_uint(a.y) & 0xFFFFE000u);
        l.z = a.z - __uint_as_float(__float_as_uint(a.z) & 0xFFFFE000u);
        l.w = a.w - __uint_as_float(__float_as_uint(a.w) & 0xFFFFE000u);
        sts4(hi + TC_STAGE_BYTES + off, l);
      }
      fence_proxy_async();                              // generic-proxy writes -> the tensor core's reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.split_bar[stage]);
      // a chunk is drained one slab into the next one (its MMAs have had time to finish: no bubble)
      if (drained < n_chunks - 1 && it >= (drained + 1) * X3_CHUNK) drain(drained++);
    }
    while (drained < n_chunks) drain(drained++);

    // ---- fused epilogue on the running sums (same code as tc_epilogue, accumulators already in registers) ----
    const Group e = ctx.g;
    const int m = m0 + lq * 32 + lane;
    const int mode = e.ksplit > 1 ? TILE_PARTIAL : TILE_FINAL;
    float* const obase = mode == TILE_PARTIAL ? e.partial + (size_t)split * e.M * e.N : e.C;
    const int ldo = mode == TILE_PARTIAL ? e.N : e.ldc;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int nb = n0 + (chalf * 2 + c) * 32;
      if (m < e.M && nb < e.N) {
        float* orow = obase + (size_t)m * ldo + nb;
        const int nvalid = min(32, e.N - nb);
        if (mode == TILE_FINAL) {
          TA3N_EPI_DISPATCH(e.flags, { epilogue_row32<EPI_F>(e, m, nb, nvalid, sum[c]); })
        }
        if (nb + 32 <= e.N && ((reinterpret_cast<uintptr_t>(orow) & 15u) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(orow + j) = make_float4(sum[c][j], sum[c][j + 1], sum[c][j + 2], sum[c][j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < e.N) orow[j] = sum[c][j];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, X3_TMEM_COLS);
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

// Per-device facts and one-time kernel configuration.  cudaFuncSetAttribute is per DEVICE (a process may drive
// several GPUs, e.g. one host thread per replica as nn.DataParallel does -- main.py:79), so the bookkeeping is
// keyed by device ordinal and guarded by a mutex.
struct DeviceInfo {
  int sm_count = 0;
  bool configured[8] = {false, false, false, false, false, false, false, false};
  bool step_configured = false;      // step_kernel.cuh kernels (ta3n_api.cu)
  bool x3_configured[4] = {false, false, false, false};      // seg_gemm_tc_x3_kernel, per operand layout
};
inline std::mutex& device_mu() {
  static std::mutex mu;
  return mu;
}
inline DeviceInfo* device_info() {      // call with device_mu() held
  static std::map<int, DeviceInfo> devs;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  DeviceInfo& d = devs[dev];
  if (d.sm_count == 0) {
    if (cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || d.sm_count <= 0)
      d.sm_count = 148;
  }
  return &d;
}
inline int device_sm_count() {
  std::lock_guard<std::mutex> lock(device_mu());
  DeviceInfo* d = device_info();
  return d ? d->sm_count : 148;
}

struct MapKey {
  const void* ptr;
  long inner, outer, ld;
  int box_inner, box_outer;
  int atom32;   // 1: SWIZZLE_128B_ATOM_32B (MN-major operands), 0: SWIZZLE_128B
  int rank3;    // 1: MN-major operand as {32 floats, outer rows, inner/32 groups}, box {32, box_outer, 4}
  int raw;      // 1: FLOAT32 (the raw words, precise kernel); 0: TFLOAT32 (TMA rounds to tf32)
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
  // a DRIVER call: the calling thread needs a current context.  A replica thread of nn.DataParallel (main.py:79) has
  // only selected its device through the runtime so far -- bind the primary context (no-op when already bound).
  cudaFree(nullptr);
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
  // bit patterns (measured in round 1: a -7e-4 bias per GEMM with FLOAT32 maps).
  CUresult r = fn(out, k.raw ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, k.rank3 ? 3 : 2,
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
  if (g.seg_count > kMaxSegs) return false;
  for (int i = 0; i < g.seg_count; ++i) {
    const Seg& s = plan.segs[g.seg_begin + i];
    if (!tc_operand_ok(s.A, s.lda) || !tc_operand_ok(s.B, s.ldb) || s.len <= 0) return false;
  }
  return true;
}

// tensor-map keys of one segment of a group (shared by the per-launch and the step-kernel planners)
inline void tc_seg_keys(const Seg& s, const Group& g, bool a_kmaj, bool b_kmaj, bool a3d, bool b3d, MapKey* ka,
                        MapKey* kb, bool raw = false) {
  const int rw = raw ? 1 : 0;
  *ka = a_kmaj ? MapKey{s.A, s.len, g.M, s.lda, TC_BK, TC_BM, 0, 0, rw} : MapKey{s.A, g.M, s.len, s.lda, 32, TC_BK, 1, a3d ? 1 : 0, rw};
  *kb = b_kmaj ? MapKey{s.B, s.len, g.N, s.ldb, TC_BK, TC_BN, 0, 0, rw} : MapKey{s.B, g.N, s.len, s.ldb, 32, TC_BK, 1, b3d ? 1 : 0, rw};
}

template <bool A_KMAJ, bool B_KMAJ, int STAGES>
inline int tc_launch_stages(const GemmTable& tab, const TcMaps& maps, const TcSegMaps& sm, cudaStream_t stream,
                            const char* label) {
  constexpr int slot = (A_KMAJ ? 0 : 1) + (B_KMAJ ? 0 : 2) + (STAGES > 3 ? 4 : 0);
  int first_wave = 148;
  {
    std::lock_guard<std::mutex> lock(device_mu());
    DeviceInfo* d = device_info();
    if (!d) return fail(TA3N_ERR_CUDA, "cudaGetDevice failed");
    if (!d->configured[slot]) {
      TA3N_CUDA(cudaFuncSetAttribute(seg_gemm_tc_kernel<A_KMAJ, B_KMAJ, STAGES>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(STAGES)));
      d->configured[slot] = true;
    }
    first_wave = d->sm_count;
  }
  pre_launch(label, stream);
  launch_kernel(seg_gemm_tc_kernel<A_KMAJ, B_KMAJ, STAGES>, tab.total_tiles, tc_threads(STAGES),
                tc_smem_bytes(STAGES), stream, tab, maps, sm, first_wave);
  return after_launch();
}

// the precise kernel, per operand layout
template <bool A_KMAJ, bool B_KMAJ>
inline int tc_launch_x3(const GemmTable& tab, const TcMaps& maps, const TcSegMaps& sm, cudaStream_t stream, const char* label) {
  constexpr int slot = (A_KMAJ ? 0 : 1) + (B_KMAJ ? 0 : 2);
  {
    std::lock_guard<std::mutex> lock(device_mu());
    DeviceInfo* d = device_info();
    if (!d) return fail(TA3N_ERR_CUDA, "cudaGetDevice failed");
    if (!d->x3_configured[slot]) {
      TA3N_CUDA(cudaFuncSetAttribute(seg_gemm_tc_x3_kernel<A_KMAJ, B_KMAJ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     x3_smem_bytes()));
      d->x3_configured[slot] = true;
    }
  }
  pre_launch(label, stream);
  launch_kernel(seg_gemm_tc_x3_kernel<A_KMAJ, B_KMAJ>, tab.total_tiles, X3_THREADS, x3_smem_bytes(), stream, tab, maps, sm);
  return after_launch();
}

template <bool A_KMAJ, bool B_KMAJ>
inline int tc_launch_one(const GemmTable& tab, const TcMaps& maps, const TcSegMaps& sm, cudaStream_t stream,
                         const char* label) {
  // 6 stages / one CTA per SM for grids within a wave, 3 stages / two CTAs per SM beyond
  return tab.total_tiles <= device_sm_count() ? tc_launch_stages<A_KMAJ, B_KMAJ, 6>(tab, maps, sm, stream, label)
                                              : tc_launch_stages<A_KMAJ, B_KMAJ, 3>(tab, maps, sm, stream, label);
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

// rank-3 maps for MN-major operands whose MN extent is a multiple of 32 in every group of the plan
inline void tc_rank3_flags(const GemmPlan& plan, bool* a3d, bool* b3d) {
  *a3d = !plan.a_kmaj;
  *b3d = !plan.b_kmaj;
  for (const Group& g : plan.groups) {
    if (g.M % 32 != 0) *a3d = false;
    if (g.N % 32 != 0) *b3d = false;
  }
}

// Launch `plan` (all groups eligible) on the tcgen05 engine.
inline int launch_tc(const GemmPlan& plan_in, cudaStream_t stream, bool precise = false) {
  // longest-K groups first (see the tile remap in the kernel): LPT-style balance of the tensor pipe
  std::vector<int> order(plan_in.groups.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    return plan_in.k_total(plan_in.groups[a]) / plan_in.groups[a].ksplit >
           plan_in.k_total(plan_in.groups[b]) / plan_in.groups[b].ksplit;
  });
  const GemmPlan plan = sub_plan(plan_in, order);
  bool a3d, b3d;
  tc_rank3_flags(plan, &a3d, &b3d);
  size_t gi = 0;
  while (gi < plan.groups.size()) {
    GemmTable tab;
    TcMaps maps;
    TcSegMaps sm;
    memset(&tab, 0, sizeof(int) * 4);
    memset(&sm, 0, sizeof(sm));
    std::map<MapKey, int> local;
    int ng = 0, ns = 0, tiles = 0, nmaps = 0;
    bool any_split = false;
    while (gi < plan.groups.size() && ng < kMaxGroups) {
      const Group& src = plan.groups[gi];
      if (ns + src.seg_count > kMaxSegs) break;
      // tensor maps of this group's segments (deduplicated within the launch)
      std::vector<std::pair<MapKey, MapKey>> keys;
      int fresh = 0;
      std::map<MapKey, int> trial = local;
      for (int i = 0; i < src.seg_count; ++i) {
        MapKey ka, kb;
        tc_seg_keys(plan.segs[src.seg_begin + i], src, plan.a_kmaj, plan.b_kmaj, a3d, b3d, &ka, &kb, precise);
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
      any_split |= g.ksplit > 1;
      tab.g[ng++] = g;
      ++gi;
    }
    tab.n_groups = ng;
    tab.total_tiles = tiles;
    tab.pad_ = (a3d ? 1 : 0) | (b3d ? 2 : 0);
    if (tiles > 0) {
      if (precise && plan.a_kmaj && plan.b_kmaj)
        TA3N_TRY((tc_launch_x3<true, true>(tab, maps, sm, stream, plan.label)));
      else if (precise && plan.a_kmaj && !plan.b_kmaj)
        TA3N_TRY((tc_launch_x3<true, false>(tab, maps, sm, stream, plan.label)));
      else if (precise && !plan.a_kmaj && !plan.b_kmaj)
        TA3N_TRY((tc_launch_x3<false, false>(tab, maps, sm, stream, plan.label)));
      else if (precise)
        TA3N_TRY((tc_launch_x3<false, true>(tab, maps, sm, stream, plan.label)));
      else if (plan.a_kmaj && plan.b_kmaj)
        TA3N_TRY((tc_launch_one<true, true>(tab, maps, sm, stream, plan.label)));
      else if (plan.a_kmaj && !plan.b_kmaj)
        TA3N_TRY((tc_launch_one<true, false>(tab, maps, sm, stream, plan.label)));
      else if (!plan.a_kmaj && !plan.b_kmaj)
        TA3N_TRY((tc_launch_one<false, false>(tab, maps, sm, stream, plan.label)));
      else
        TA3N_TRY((tc_launch_one<false, true>(tab, maps, sm, stream, plan.label)));
      if (any_split) {
        // vectorised reduce when every split group allows it (the forward layers), blocks for split groups only
        SplitGroups sgs;
        sgs.n = 0;
        bool vec = true;
        size_t mx4 = 0;
        for (int i = 0; i < ng; ++i) {
          const Group& g = tab.g[i];
          if (g.ksplit <= 1) continue;
          sgs.idx[sgs.n++] = (unsigned char)i;
          const bool ok = g.N % 4 == 0 && g.ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(g.C) & 15u) == 0 &&
                          (reinterpret_cast<uintptr_t>(g.partial) & 15u) == 0 &&
                          (g.flags & ~(EPI_BIAS | EPI_RELU | EPI_DROP_RNG)) == 0 && g.ksplit <= 8;
          vec = vec && ok;
          mx4 = std::max(mx4, (size_t)g.M * g.N / 4);
        }
        if (vec && sgs.n > 0) {
          dim3 grid((unsigned)std::min<size_t>((mx4 + 255) / 256, 1024), sgs.n);
          pre_launch("splitk_reduce", stream);
          launch_kernel(splitk_reduce_v4_kernel, grid, 256, 0, stream, tab, sgs);
        } else {
          dim3 grid(splitk_reduce_blocks(tab), ng);
          pre_launch("splitk_reduce", stream);
          launch_kernel(splitk_reduce_kernel, grid, 256, 0, stream, tab);
        }
        TA3N_TRY(after_launch());
      }
    }
  }
  return TA3N_OK;
}

// ---- balanced split-K for the precise forward launches ------------------------------------------------------
// The precise kernel holds one CTA per SM and is bound by shared-memory bandwidth (~0.9 us per K slab), so a launch
// costs what its longest CTA queue costs: the shared layer has 80 tiles of 64 slabs for 148 SMs, the forward batch 80
// tiles of 16 slabs next to tiles of up to 80.  Split factors per group are chosen by simulating the greedy (LPT)
// assignment of the resulting tasks to the SMs for a few target task lengths; partials go to the caller's scratch
// (ta3n_set_forward_scratch) and a fixed-order reduce pass applies the epilogue.  Deterministic.
struct ForwardScratch {
  void* ptr = nullptr;
  size_t bytes = 0;
};
inline ForwardScratch& forward_scratch() {
  static thread_local ForwardScratch s;
  return s;
}
inline int x3_plan_slabs(const GemmPlan& plan, const Group& g) {
  int n = 0;
  for (int k = 0; k < g.seg_count; ++k) n += (plan.segs[g.seg_begin + k].len + TC_BK - 1) / TC_BK;
  return n;
}
inline double x3_makespan(const GemmPlan& plan, const std::vector<int>& ks, int sms) {
  std::vector<double> tasks;
  for (size_t gi = 0; gi < plan.groups.size(); ++gi) {
    const Group& g = plan.groups[gi];
    const int tiles = ((g.M + TC_BM - 1) / TC_BM) * ((g.N + TC_BN - 1) / TC_BN);
    const double len = (double)x3_plan_slabs(plan, g) / ks[gi] + 4.0;      // + prologue / epilogue, in slab units
    for (int t = 0; t < tiles * ks[gi]; ++t) tasks.push_back(len);
  }
  std::sort(tasks.begin(), tasks.end(), std::greater<double>());
  std::vector<double> load(sms, 0.0);
  for (double t : tasks) *std::min_element(load.begin(), load.end()) += t;
  return *std::max_element(load.begin(), load.end());
}
inline void plan_splitk_balanced(GemmPlan& plan, Arena* arena, int sms) {
  if (!arena) return;
  double total = 0;
  for (const Group& g : plan.groups)
    total += (double)x3_plan_slabs(plan, g) * ((g.M + TC_BM - 1) / TC_BM) * ((g.N + TC_BN - 1) / TC_BN);
  std::vector<int> best(plan.groups.size(), 1);
  double best_cost = x3_makespan(plan, best, sms);
  for (double c : {1.1, 1.25, 1.5, 2.0, -2.0, -3.0, -4.0}) {      // negative: the same factor for every group
    const double target = std::max(8.0, total / sms * c);
    std::vector<int> ks(plan.groups.size(), 1);
    for (size_t gi = 0; gi < plan.groups.size(); ++gi) {
      const int slabs = x3_plan_slabs(plan, plan.groups[gi]);
      int k = c > 0 ? (int)std::ceil(slabs / target) : (int)(-c);
      k = std::max(1, std::min(k, 8));
      while (k > 1 && slabs / k < 8) --k;
      ks[gi] = k;
    }
    bool any = false;
    for (int k : ks) any |= k > 1;
    const double cost = x3_makespan(plan, ks, sms) + (any ? 8.0 : 0.0);      // the reduce pass
    if (cost < best_cost * 0.92) {
      best_cost = cost;
      best = ks;
    }
  }
  for (size_t gi = 0; gi < plan.groups.size(); ++gi) {
    Group& g = plan.groups[gi];
    if (best[gi] < 2) continue;
    float* p = arena->floats((size_t)best[gi] * g.M * g.N);
    if (!p) continue;                      // scratch too small: stay unsplit (still correct)
    g.ksplit = best[gi];
    g.partial = p;
  }
}

// TA3N_X3_DGRAD=1: the data-gradient GEMMs of the x3 engine at fp32 grade as well.  Measured at cfg2 (profiles/
// r2_x3_ab.txt): +36 us per step (340 -> 376) for TRN bias gradients at 2e-6 instead of 2.4e-4 and no change of the
// worst tensors (shared-layer gradients 1.24e-3 vs 1.29e-3 unpinned: those are set by ReLU-pattern differences and by
// the tf32 weight-gradient GEMM itself) -> off by default.
inline bool x3_dgrad_enabled() {
  static const bool on = []() {
    const char* e = getenv("TA3N_X3_DGRAD");
    return e && e[0] == '1';
  }();
  return on;
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
    // the precise kernel for the forward layers of the x3 engine (K-major operands); everything else plain tf32
    const bool precise = engine == TA3N_GEMM_TF32X3_TCGEN05 && (plan.precise || (plan.precise_dgrad && x3_dgrad_enabled()));
    std::vector<int> tc_idx, simt_idx;
    // x3 engine: the small weight-gradient GEMMs (the 256 x 256 layers of the video / relation discriminators,
    // <= 0.3 GFLOP each) run on the precise kernel.  Near the adversarial equilibrium their source and target
    // halves cancel, which amplifies the 3e-4 of a tf32 product to 4e-3 of the net gradient (cfg3, measured).
    // (On the exact SIMT engine the same five groups cost 64 us; as one 20-tile launch of the x3 kernel ~15 us.)
    const bool small_exact = engine == TA3N_GEMM_TF32X3_TCGEN05 && !plan.a_kmaj && !plan.b_kmaj && !precise;
    std::vector<int> fine_idx;
    for (int i = 0; i < (int)plan.groups.size(); ++i) {
      const Group& g = plan.groups[i];
      const bool tiny = small_exact && (long)g.M * g.N <= 256L * 256L && 2.0 * g.M * g.N * (double)plan.k_total(g) <= 3e8;
      if (!tc_group_ok(plan, g))
        simt_idx.push_back(i);
      else
        (tiny ? fine_idx : tc_idx).push_back(i);
    }
    if (!fine_idx.empty()) {
      GemmPlan fine = sub_plan(plan, fine_idx);
      fine.label = "wgrad_small_x3";
      TA3N_TRY(launch_tc(fine, stream, true));
    }
    if (!tc_idx.empty()) {
      GemmPlan tc = sub_plan(plan, tc_idx);
      // only the forward plans: they run on ONE stream, in order, so they can share the scratch; precise data-gradient
      // plans (TA3N_X3_DGRAD=1) may run on forked streams (parallel_branches) and stay unsplit
      if (precise && plan.precise && splitk_arena == nullptr && forward_scratch().ptr != nullptr) {
        Arena scratch(forward_scratch().ptr, forward_scratch().bytes);
        plan_splitk_balanced(tc, &scratch, device_sm_count());
      } else {
        plan_splitk(tc, splitk_arena, TC_BM, TC_BN, TC_BK, 4);
      }
      TA3N_TRY(launch_tc(tc, stream, precise));
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
