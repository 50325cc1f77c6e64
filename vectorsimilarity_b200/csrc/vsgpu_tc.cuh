// tcgen05 / TMA / mbarrier PTX wrappers shared by the tensor-core kernels (vsgpu_tensor.cu: kind::f16,
// vsgpu_tensor_i8.cu: kind::i8), plus the driver entry point for tensor-map encoding.
#pragma once
#include "vsgpu_internal.cuh"
#include <cuda.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <utility>
#include <vector>

namespace vsgpu {

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trap (launch error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr)
                 : "memory");
}
// 32 floats from shared memory as 8 x LDS.128 (the operand structs are reached through an integer-aligned
// pointer, so plain indexing compiles to generic LD.E — one dependent generic load per element)
__device__ __forceinline__ void lds_f32x32(uint32_t saddr, float (&t)[32]) {
#pragma unroll
    for (int v = 0; v < 8; v++)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(t[4 * v]), "=f"(t[4 * v + 1]), "=f"(t[4 * v + 2]), "=f"(t[4 * v + 3])
                     : "r"(saddr + 16u * v));
}
// Appends the hits of one 32-row x 32-column accumulator chunk to the per-query candidate lists; the whole warp calls it.
// `hit` is this lane's (row's) mask over the chunk's 32 columns (queries q0 .. q0+31). The masks are transposed with
// ballots so that lane j owns column j and reserves that column's slots with ONE atomic — 32 different counters in a
// single instruction, one round trip per chunk. (One atomic per hit serialised ~30 dependent round trips per chunk at
// the 1-8 % pass rates of the early phases and left the epilogue, not the MMAs, as the critical path.)
__device__ __forceinline__ void warp_append_hits(uint32_t hit, uint32_t q0, uint32_t row, const uint32_t (&r)[32],
                                                 uint32_t *__restrict__ cnt, uint2 *__restrict__ cand, int lane, uint32_t CAP) {
    const uint32_t any = __reduce_or_sync(0xffffffffu, hit);
    if (!any) return;
    uint32_t mine = 0; // rows of this warp that hit column `lane`
#pragma unroll
    for (int j = 0; j < 32; j++) {
        if ((any >> j) & 1u) { // warp-uniform
            const uint32_t b = __ballot_sync(0xffffffffu, (hit >> j) & 1u);
            if (lane == j) mine = b;
        }
    }
    uint32_t base = 0;
    if (mine) base = atomicAdd(&cnt[q0 + (uint32_t)lane], (uint32_t)__popc(mine));
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < 32; j++) {
        if ((any >> j) & 1u) {
            const uint32_t b = __shfl_sync(0xffffffffu, mine, j), bs = __shfl_sync(0xffffffffu, base, j);
            if ((hit >> j) & 1u) {
                const uint32_t slot = bs + (uint32_t)__popc(b & lt);
                if (slot < CAP) cand[(size_t)(q0 + (uint32_t)j) * CAP + slot] = make_uint2(row, r[j]);
            }
        }
    }
}
// Sparse hits (the long late phases: a hit in 1-10 % of the chunks, almost always a single one): the ballot-transposed append
// above stalls the warp for the full round trip of its returning atomic — measured at ~2 chunk times per call, which made a
// phase with hits in 9 % of its chunks 20 % slower than one without (profiles/r2_launches_i8_shard.md). Here a lane with a hit
// reserves its slot with its own atomic and only RECORDS (slot, query, row, value); the store that needs the slot happens
// when the record is recycled two sparse appends later (or at the end of the kernel), so the round trip overlaps the following
// chunks' loads and compares. Chunks where some lane has more than one hit take the dense path.
struct DeferredHits {
    uint32_t slot0, q0, row0, val0, slot1, q1, row1, val1;
    uint32_t which; // warp-uniform: the record the next sparse append recycles
    __device__ __forceinline__ void init() {
        q0 = q1 = 0xffffffffu;
        slot0 = slot1 = row0 = row1 = val0 = val1 = 0;
        which = 0;
    }
    __device__ __forceinline__ void flush(uint2 *__restrict__ cand, uint32_t CAP) {
        if (q0 != 0xffffffffu && slot0 < CAP) cand[(size_t)q0 * CAP + slot0] = make_uint2(row0, val0);
        if (q1 != 0xffffffffu && slot1 < CAP) cand[(size_t)q1 * CAP + slot1] = make_uint2(row1, val1);
        q0 = q1 = 0xffffffffu;
    }
};
__device__ __forceinline__ void warp_append_hits_deferred(uint32_t hit, uint32_t qbase, uint32_t row, const uint32_t (&r)[32],
                                                          uint32_t *__restrict__ cnt, uint2 *__restrict__ cand, int lane, uint32_t CAP,
                                                          DeferredHits &d) {
    const uint32_t any = __reduce_or_sync(0xffffffffu, hit);
    if (!any) return;
    if (__any_sync(0xffffffffu, (hit & (hit - 1u)) != 0u)) { // some row hit several queries of this chunk: dense path
        warp_append_hits(hit, qbase, row, r, cnt, cand, lane, CAP);
        return;
    }
    // the value of the single hit column, without indexing the register array dynamically
    uint32_t v16[16], v8[8], v4[4], v2[2];
#pragma unroll
    for (int i = 0; i < 16; i++) v16[i] = (hit & 0xaaaaaaaau) ? r[2 * i + 1] : r[2 * i];
#pragma unroll
    for (int i = 0; i < 8; i++) v8[i] = (hit & 0xccccccccu) ? v16[2 * i + 1] : v16[2 * i];
#pragma unroll
    for (int i = 0; i < 4; i++) v4[i] = (hit & 0xf0f0f0f0u) ? v8[2 * i + 1] : v8[2 * i];
#pragma unroll
    for (int i = 0; i < 2; i++) v2[i] = (hit & 0xff00ff00u) ? v4[2 * i + 1] : v4[2 * i];
    const uint32_t val = (hit & 0xffff0000u) ? v2[1] : v2[0];
    const uint32_t q = qbase + (uint32_t)(__ffs((int)hit) - 1);
    if (d.which == 0) { // warp-uniform
        if (d.q0 != 0xffffffffu && d.slot0 < CAP) cand[(size_t)d.q0 * CAP + d.slot0] = make_uint2(d.row0, d.val0);
        d.q0 = 0xffffffffu;
        if (hit) {
            d.slot0 = atomicAdd(&cnt[q], 1u);
            d.q0 = q;
            d.row0 = row;
            d.val0 = val;
        }
    } else {
        if (d.q1 != 0xffffffffu && d.slot1 < CAP) cand[(size_t)d.q1 * CAP + d.slot1] = make_uint2(d.row1, d.val1);
        d.q1 = 0xffffffffu;
        if (hit) {
            d.slot1 = atomicAdd(&cnt[q], 1u);
            d.q1 = q;
            d.row1 = row;
            d.val1 = val;
        }
    }
    d.which ^= 1u;
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// The same wait, naming the registers a preceding tcgen05.ld fills as read-write operands: the compiler may otherwise move
// plain reads of those registers above the wait (nothing else ties them to it) — seen as stale accumulators once two
// loads were kept in flight.
__device__ __forceinline__ void tc_wait_ld(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// K-major, SWIZZLE_128B operand descriptor: rows at a 128-byte pitch, 8-row groups 1024 bytes apart
// (SBO = 64 x 16 B), LBO unused (1), descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fff) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}


// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster run one M = 256 MMA, each staging its own 128 rows of A and
// its half of B; the leader (cluster rank 0) issues the MMAs and owns the "operands landed" barriers ------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_i8_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
// arrives (once the MMAs issued so far have completed) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tc_encode_fn() {
    // function-local static: initialised exactly once even when several shard threads arrive together
    static const EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        EncodeTiledFn f = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            f = (EncodeTiledFn)p;
        cudaGetLastError();
        return f;
    }();
    return fn;
}

// Row ranges of the filtered GEMM's phases: [0, s0) unfiltered, then edges in a common ratio <= growth so that a phase
// admits ~ (growth - 1) * k rows per query (growth is what the per-phase candidate buffer affords). The fewest phases
// with that property, the last one ending exactly at n: every launch carries a fixed ~50-80 us (pipeline fill, tail,
// the merge that follows), so a short trailing phase is the most expensive way to finish.
//
// Sharded callers that exchange their bounds between the phases (vsgpu_topk_device_begin / _next / _finish, `world` shards):
// the exchanged bound is the shards' ceil(k / world)-th best score, not the k-th, so a phase of G times the rows seen
// admits ~ G * k / world rows per query and the same buffer affords a `world` times larger growth: 3 phases instead of 5
// at 1.25 M rows per shard. All shards must run the same number of phases (`rounds`, from topk_rounds on a row count
// every shard agrees on), whatever their own row count.
inline double phase_growth(size_t k, size_t cand_cap, unsigned world) {
    if (world <= 1) return std::max(3.0, std::min(8.0, (double)cand_cap / (2.5 * (double)k)));
    const double m = (double)((k + world - 1) / world);
    return std::max(3.0, std::min(64.0, (double)cand_cap / (4.0 * m))); // the min over the shards of an m-th order statistic is looser
}
inline size_t phase_first(size_t k, size_t cand_cap, size_t bm) {
    size_t s0 = std::max<size_t>(bm, std::min<size_t>(cand_cap, 2048) / bm * bm);
    s0 = std::max(s0, (std::min<size_t>(2 * k, cand_cap) + bm - 1) / bm * bm);
    if (const char *e = getenv("VSGPU_PHASE_S0")) { // experiments only
        const size_t v = (size_t)atol(e) / bm * bm;
        if (v >= s0 && v <= cand_cap) s0 = v;
    }
    return s0;
}
inline size_t topk_rounds(size_t n, size_t k, size_t cand_cap, size_t bm, unsigned world) {
    const size_t s0 = phase_first(k, cand_cap, bm);
    size_t np = 1;
    if (n > s0) np += (size_t)std::ceil(std::log((double)n / (double)s0) / std::log(phase_growth(k, cand_cap, world)) - 1e-9);
    return np;
}
// rounds == 0: as many phases as this n needs. rounds > 0: exactly that many (trailing ones may be empty: first == second)
inline std::vector<std::pair<uint32_t, uint32_t>> make_phases(size_t n, size_t k, size_t cand_cap, size_t bm, unsigned world = 1,
                                                              size_t rounds = 0) {
    std::vector<std::pair<uint32_t, uint32_t>> phases;
    const size_t s0 = phase_first(k, cand_cap, bm);
    const size_t np = rounds ? rounds : topk_rounds(n, k, cand_cap, bm, world);
    const double ratio = np > 1 && n > s0 ? std::pow((double)n / (double)s0, 1.0 / (double)(np - 1)) : 1.0;
    size_t a = 0;
    double edge = (double)std::min(n, s0);
    for (size_t p = 0; p < np; p++) {
        const size_t b = a >= n ? n : p + 1 == np ? n : std::min(n, std::max((size_t)edge / bm * bm, a + bm));
        phases.emplace_back((uint32_t)a, (uint32_t)b);
        a = b;
        edge *= ratio;
    }
    if (a < n) phases.emplace_back((uint32_t)a, (uint32_t)n);
    return phases;
}

} // namespace vsgpu
