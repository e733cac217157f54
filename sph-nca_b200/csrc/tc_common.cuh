// Thin inline-PTX wrappers for the Blackwell tensor path used by the fused MLP kernels:
// tcgen05.mma (kind::tf32, A from smem or TMEM), TMEM alloc/ld/st, mbarrier, proxy fences.
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor"
// tables (same fields CUTLASS's cute/arch/mma_sm100_desc.hpp encodes).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sphnca {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- operand layout in shared memory: K-major, no swizzle ("interleave") -----------------------
// A tile of R rows x K tf32 (fp32 containers) is stored as 16-byte chunks (4 consecutive k of one
// row); chunk (row m, c = k/4) lives at byte offset (c*R + m)*16.  An 8-row x 16 B core matrix is
// therefore 128 contiguous bytes; the next core matrix along M/N is +128 B (SBO) and the next
// 16-byte K chunk is +R*16 B (LBO).  A thread that owns row m writes chunk c with one conflict-free
// STS.128 (consecutive rows -> consecutive 16 B).
__host__ __device__ constexpr uint32_t chunk_off(uint32_t rows, uint32_t m, uint32_t c) { return (c * rows + m) * 16u; }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);                 // start address, bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;       // leading-dimension byte offset, [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;       // stride-dimension byte offset, [32,46)
    d |= (uint64_t)1 << 46;                                  // descriptor version (Blackwell), [46,48)
    return d;                                                // base_offset 0, lbo_mode 0, SWIZZLE_NONE (0) in [61,64)
}

// The descriptor of the same tile `bytes` further on (bytes % 16 == 0, address stays below 256 KB: no carry out of the 14-bit field).
// Build the base descriptors ONCE, before the tile loop, and advance them with this: base + constant, instead of re-deriving
// every descriptor from its address (shift, mask, or) in front of its MMA.  10 % fewer instructions in the forward kernel,
// -2.5 % kernel time; the MMA issue rate itself was not what limited the kernels.
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }

// kind::tf32, D = fp32, A and B K-major, dense
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t M, uint32_t N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// ---- warp-uniform issue ------------------------------------------------------------------------
// The whole MMA warp runs the issue loop and the instruction itself is predicated on ONE elected lane.
// Operands derived from warp-uniform values then stay in uniform registers (UIADD3.64 + UTCHMMA, ~3 SASS
// instructions per MMA); issuing from `if (lane == 0)` instead makes ptxas move every descriptor through
// R2UR inside a BRA.U.ANY convergence loop (~16 instructions) and the issuing thread, not the tensor core,
// becomes the floor: 72-90 cycles per M128 x N<=128 x K8 MMA vs the tensor pipe's own N/2 cycles with this form
// (csrc/probes/tc_timing3.cu vs tc_timing4.cu / tc_timing5.cu).
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void mma_tf32_ss_e(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate, uint32_t go) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate), "r"(go) : "memory");
}
__device__ __forceinline__ void mma_tf32_ts_e(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate, uint32_t go) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate), "r"(go) : "memory");
}
__device__ __forceinline__ void mma_commit_e(uint64_t* bar, uint32_t go) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" :: "r"(smem_u32(bar)), "r"(go) : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// ---- TMEM ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp gets columns [col, col+16) of lane
// (lane_base + i).  taddr = (lane << 16) | column.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
        :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(smem_u32(bar)) : "memory");
}
// (try_wait with an explicit suspend-time hint was tried to keep polling warps off the issue slots: no gain in the
// MLP kernels and a slower wake-up of the TMA-staged sweep, 0.31 vs 0.27 ms -- the default time limit stays.)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// ---- TMA bulk copy (1-D): global -> shared, completion counted in bytes on an mbarrier ------------
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug traps instead of hanging the GPU (try_wait itself sleeps in HW).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // not unrolled: the compiler otherwise emits ~32 try_wait/branch pairs per call site (768 in the forward MLP kernel), and
    // ncu showed 13 % of that kernel's warp stalls as "no instructions" (instruction-cache misses)
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 24); ++it)
        if (mbar_try_wait(bar, parity)) return;
    asm volatile("trap;");
}

// Same, with a short sleep between polls: a warp that spins on try_wait keeps issuing shared-memory
// instructions and competes with the warps doing real LDS/STS/LDGSTS work for the SM's memory-instruction queue.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t ns = 32) {
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        if (mbar_try_wait(bar, parity)) return;
        __nanosleep(ns);
    }
    asm volatile("trap;");
}

// ---- 3xTF32 split ----------------------------------------------------------------------------------
// x = hi + lo with hi = x rounded to tf32 (10 explicit mantissa bits, round-to-nearest) and lo the
// fp32 remainder (itself fed to the tensor core, which reads its top 19 bits).  hi*hi + hi*lo + lo*hi
// reproduces the fp32 product to ~2^-21 relative.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    // round-to-nearest (ties away) to 10 mantissa bits on the integer ALU: add half an ulp of the kept field to the
    // magnitude and clear the 13 dropped bits -- the value `cvt.rna.tf32.f32` gives for finite inputs, in 2 full-rate
    // instructions instead of the ~6-instruction sequence ptxas emits for the cvt (the epilogues are issue-bound)
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    lo = x - hi;
}

// Weight staging at kernel start: U independent global loads per thread are issued before the first of them is used, so a
// CTA pays the L2 latency total/(threads*U) times instead of total/threads times (the plain loop cost ~13 us per launch for
// the 84 KB of weights -- visible now that a launch of the MLP kernels is only ~100 us).
template <int U, typename L, typename S>
__device__ __forceinline__ void stage_batched(int total, int tid, int threads, L load, S store) {
    for (int i0 = tid; i0 < total; i0 += threads * U) {
        float w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int i = i0 + u * threads; w[u] = i < total ? load(i) : 0.f; }
#pragma unroll
        for (int u = 0; u < U; ++u) { const int i = i0 + u * threads; if (i < total) store(i, w[u]); }
    }
}

// Weight staging into the K-major tile layout (chunk_off) as tf32 hi / lo, one 16-byte K chunk (4 K-consecutive values of
// one tile row) per thread and step.  Lane mapping inside a warp: 8 consecutive rows x 4 consecutive chunks -- each quarter
// warp writes 128 contiguous bytes (no bank conflict) and, where K is the fast dimension in global memory, a warp-level load
// covers 64 contiguous bytes of 8 rows.  stage_batched's per-element form with consecutive threads along K put 32 lanes on 4
// banks (8-way conflict on both stores): 11,900 cycles per launch for the forward kernel's 84 KB of weights -- a third of an
// empty launch.  Tiles need rows % 8 == 0 and chunks % 4 == 0.
// lane mapping of the idx-th chunk of a tile with `rgroups` groups of 8 rows (see above)
__device__ __forceinline__ void stage_map(int idx, int rgroups, int& row, int& chunk) {
    const int blk = idx >> 5, l = idx & 31, q = blk / rgroups, g = blk - q * rgroups;
    row = 8 * g + (l & 7); chunk = 4 * q + (l >> 3);
}
// get4(idx) loads chunk idx of a tile (U loads of a thread in flight at once); put(idx, hi, lo) stores its tf32 hi / lo parts.
// One call per tile with U = chunks per thread: enumerating ALL tiles of a launch in one index space (10 loads in flight, one
// round trip) was measured SLOWER (9,200 vs 7,700 cycles for the forward kernel's prologue) -- the unrolled code is twice as
// long and a cold launch fetches it line by line.
template <int U, typename G, typename P>
__device__ __forceinline__ void stage_chunks(int total, int tid, int threads, G get4, P put) {
    for (int u0 = tid; u0 < total; u0 += threads * U) {
        float4 w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int idx = u0 + u * threads; if (idx < total) w[u] = get4(idx); }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = u0 + u * threads;
            if (idx < total) {
                float hi[4], lo[4];
                split_tf32(w[u].x, hi[0], lo[0]); split_tf32(w[u].y, hi[1], lo[1]);
                split_tf32(w[u].z, hi[2], lo[2]); split_tf32(w[u].w, hi[3], lo[3]);
                put(idx, make_float4(hi[0], hi[1], hi[2], hi[3]), make_float4(lo[0], lo[1], lo[2], lo[3]));
            }
        }
    }
}

}  // namespace tc

// development aid: per-role clock64 timestamps of CTA 0, 16 slots per tile/chunk iteration.  One buffer
// per translation unit (no relocatable device code); each kernel file exports its own reader.
static __device__ long long g_timeline[64 * 16];
// (compiled in with `make EXTRA=-DSPHNCA_TIMELINE`; off by default: the stamps cost instructions and code size in every role)
#ifdef SPHNCA_TIMELINE
#define SPHNCA_TL(slot) do { if (blockIdx.x == 0 && it < 63) sphnca::g_timeline[it * 16 + (slot)] = clock64(); } while (0)
// launch phases of CTA 0 (row 63 of the table): 0 entry, 1 weights staged, 2 barriers + TMEM ready, 3 all roles done
#define SPHNCA_TL_PHASE(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) sphnca::g_timeline[63 * 16 + (k)] = clock64(); } while (0)
#else
#define SPHNCA_TL(slot) do { } while (0)
#define SPHNCA_TL_PHASE(k) do { } while (0)
#endif
#define SPHNCA_TIMELINE_READER(name)                                                              \
    extern "C" int32_t name(long long* out_host, int32_t n) {                                     \
        if (n > 64 * 16) n = 64 * 16;                                                             \
        return cudaMemcpyFromSymbol(out_host, sphnca::g_timeline, sizeof(long long) * n) == cudaSuccess ? 0 : -2; \
    }
}  // namespace sphnca
