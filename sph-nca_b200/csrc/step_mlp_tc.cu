// NCA step forward, split form: this kernel is the MLP + update half (tcgen05), fed from HBM.
//
//   step_perceive_kernel (step.cu)   neighbour sweep only: full L1 (no shared memory), high occupancy;
//                                    writes the scaled perception Y_s [N,32] (needed by backward anyway)
//                                    and the previous-life-mask bit.
//   step_mlp_tc_kernel (this file)   loader warps pull rows [A | Y] (192 B / particle) one tile ahead into
//                                    registers and write the tf32 hi/lo MMA A tile; MMA warp and TMEM
//                                    pipeline as in step_tc.cu; 8 epilogue warps (two per TMEM lane
//                                    quarter, half of the hidden columns each) for the hidden layer and
//                                    4 finalizer warps (gates, update, row stores) one tile behind them.
//
// GEMM2 has 32 output columns, where an MMA costs ~32 cycles + N/4 (probes/tc_timing4: A from TMEM, N=32 40 cycles,
// N=64 57): W2 hi and lo are STACKED along N (64 rows), H_hi . [W_hi ; W_lo] is one MMA into D2 columns 0..31 | 32..63,
// H_lo . W_hi adds to columns 0..31 -- 2 MMAs per k-step instead of 3; the final epilogue adds the column groups.
//
// Why split: the fused kernel (step_tc.cu) needs 215 KB of shared memory, which leaves ~13 KB of L1 per SM,
// so its neighbour gathers all go to L2 and the sweep (41-58 k cycles per 128-particle tile, measured with
// the in-kernel timeline) hides the whole MLP chain.  The extra HBM traffic of the split is one 192-byte
// row read per particle.  Both paths are kept (sphnca_set_mlp_path) and produce identical tensors.
#include <type_traits>
#include "step_common.cuh"
#include "tc_common.cuh"

namespace sphnca {

using namespace tc;

namespace mt {
constexpr int TILE = 128;
constexpr int EPI_WARPS = 8, MMA_WARP = 8, LOAD_WARP0 = 9, LOAD_WARPS = 4, FIN_WARP0 = 13, FIN_WARPS = 4;
constexpr int THREADS = (FIN_WARP0 + FIN_WARPS) * 32;          // 544
constexpr int N2 = 32;
// TMEM: one 128-column region per hidden half (GEMM1 writes the pre-activations there, the epilogue converts them IN PLACE to
// H_hi), one shared H_lo region, D2 in two slots of 64 columns
constexpr uint32_t COL_DA = 0, COL_DB = 128, COL_HLO = 256, COL_D2 = 384, TMEM_COLS = 512;
enum { Y_FULL = 0, Y_EMPTY, D1_FULL, D1_FULL1, H_FULL, H_FULL1, H_EMPTY, D2_FULL, D2_FULL1, D2_EMPTY, D2_EMPTY1, FIN_FULL, FIN_FULL1, NUM_BARS };

struct Smem {
    float W1_hi[2][128 * NIN], W1_lo[2][128 * NIN];
    float W2[2][2 * N2 * 128];          // [hidden half] rows 0..31 = outputs hi, 32..63 = lo ; K = h (128)
    float Y_hi[TILE * NIN], Y_lo[TILE * NIN];
    float b1[NH], w2_last[NH], b2[NO_PAD];
    float o32[2][2][TILE];              // [slot][column-half group] partial scalar-gate sums
    float ostage[FIN_WARPS][16 * NO_PAD];  // finalizer: 16 output rows per warp on their way to a coalesced store
    uint64_t bar[NUM_BARS];
    uint32_t tmem_base;
};
__device__ __forceinline__ uint32_t fidx(uint32_t rows, uint32_t m, uint32_t k) { return chunk_off(rows, m, k >> 2) / 4 + (k & 3); }
// gates: fast exponential (2 ulp) -> |error| ~1e-7 on sigmoid / tanh, far inside the 1e-5 parity bar
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// loads whose results must not be touched before they are needed (see the epilogue's prefetch)
__device__ __forceinline__ uint32_t ldg_u8_raw(const uint8_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_f4_raw(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float fast_tanh(float x) {
    const float e = __expf(-2.f * fabsf(x));
    const float t = __fdividef(1.f - e, 1.f + e);
    return copysignf(t, x);
}
}  // namespace mt

__global__ void __launch_bounds__(mt::THREADS, 1)
step_mlp_tc_kernel(int n_total, const float* __restrict__ A_s, const float* __restrict__ Y_s,
                   const uint8_t* __restrict__ fire_s, sphnca_step_params prm, StepConsts c,
                   float* __restrict__ An_s, float* __restrict__ O_s, uint8_t* __restrict__ flags_s,
                   uint8_t* __restrict__ alive_s, uint32_t* __restrict__ relu_s, int num_tiles,
                   const int32_t* __restrict__ rows, const int32_t* __restrict__ count_ptr,
                   float* __restrict__ A_c, float* __restrict__ Y_c, int in_by_list, int fold_mask) {
    using namespace mt;
    // Fire-list mode (rows != NULL, firelist.cu): tile row j stands for sorted particle rows[j], j < *count (the count is
    // read from device memory: no host synchronisation); every listed particle fired.  O_s, relu_s and the compact copies
    // A_c, Y_c (the weight-gradient kernel's contiguous input) are indexed by j, the state / flag arrays by the particle.
    // in_by_list: A_s and Y_s ARE such compact arrays already (the fired-rows perception sweep wrote them): rows are read by j.
    if (count_ptr != nullptr) {
        n_total = __ldg(count_ptr);
        num_tiles = (n_total + TILE - 1) / TILE;
    }
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    SPHNCA_TL_PHASE(0);

    // Loader warps: lane mapping and the row loads of a tile (see the loader role below).  The loads of the CTA's FIRST tile are
    // issued here, in front of the weight staging, so that their global-memory latency runs beside it.
    const bool is_loader = warp >= LOAD_WARP0 && warp < FIN_WARP0;
    const int lw = warp - LOAD_WARP0, r8 = lane & 7, c4 = lane >> 3;
    float4 v[NIN / 4];                                        // [3 * g + {0: A chunk c4, 1: Y chunk c4, 2: Y chunk 4 + c4}], row 32 lw + 8 g + r8
    auto load_rows = [&](const int tile) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int j = tile * TILE + 32 * lw + 8 * g + r8;
            if (j < n_total) {
                const int p = (rows && !in_by_list) ? __ldg(rows + j) : j;
                const float4* a4 = reinterpret_cast<const float4*>(A_s + (size_t)p * NF);
                const float4* y4 = reinterpret_cast<const float4*>(Y_s + (size_t)p * 2 * NF);
                v[3 * g] = __ldg(a4 + c4); v[3 * g + 1] = __ldg(y4 + c4); v[3 * g + 2] = __ldg(y4 + 4 + c4);
                if (A_c != nullptr) {                         // compact copies for the backward
                    reinterpret_cast<float4*>(A_c + (size_t)j * NF)[c4] = v[3 * g];
                    float4* yc = reinterpret_cast<float4*>(Y_c + (size_t)j * 2 * NF);
                    yc[c4] = v[3 * g + 1]; yc[4 + c4] = v[3 * g + 2];
                }
            } else {
                v[3 * g] = v[3 * g + 1] = v[3 * g + 2] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    if (is_loader && (int)blockIdx.x < num_tiles) load_rows(blockIdx.x);

    const int n_mma = c.n_out < N2 ? c.n_out : N2;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        // W1 [h][k] -> W1_hi / W1_lo [half]: rows = h, K = k (contiguous in global memory: one LDG.128 per chunk)
        stage_chunks<3>(128 * (NIN / 4), tid, THREADS, [&](int idx) {
            int row, ck; stage_map(idx, 128 / 8, row, ck);
            return __ldg(reinterpret_cast<const float4*>(prm.W1 + (size_t)(half * 128 + row) * NIN) + ck);
        }, [&](int idx, float4 hi, float4 lo) {
            int row, ck; stage_map(idx, 128 / 8, row, ck);
            const uint32_t o = chunk_off(128, row, ck) / 4;
            *reinterpret_cast<float4*>(&s.W1_hi[half][o]) = hi; *reinterpret_cast<float4*>(&s.W1_lo[half][o]) = lo;
        });
        // W2 [o][h] -> W2[half]: rows 0..31 = o hi, 32..63 = o lo, K = h within the half
        stage_chunks<2>(N2 * (128 / 4), tid, THREADS, [&](int idx) {
            int row, ck; stage_map(idx, N2 / 8, row, ck);
            return row < n_mma ? __ldg(reinterpret_cast<const float4*>(prm.W2 + (size_t)row * NH + half * 128) + ck) : make_float4(0.f, 0.f, 0.f, 0.f);
        }, [&](int idx, float4 hi, float4 lo) {
            int row, ck; stage_map(idx, N2 / 8, row, ck);
            *reinterpret_cast<float4*>(&s.W2[half][chunk_off(2 * N2, row, ck) / 4]) = hi;
            *reinterpret_cast<float4*>(&s.W2[half][chunk_off(2 * N2, N2 + row, ck) / 4]) = lo;
        });
    }
    for (int i = tid; i < NH; i += THREADS) {
        s.b1[i] = __ldg(prm.b1 + i);
        s.w2_last[i] = c.n_out > N2 ? __ldg(prm.W2 + (size_t)N2 * NH + i) : 0.f;
    }
    if (tid < NO_PAD) s.b2[tid] = tid < c.n_out ? __ldg(prm.b2 + tid) : 0.f;
    SPHNCA_TL_PHASE(1);
    if (tid == 0) {
        mbar_init(&s.bar[Y_FULL], LOAD_WARPS * 32); mbar_init(&s.bar[Y_EMPTY], 1);
        mbar_init(&s.bar[D1_FULL], 1);              mbar_init(&s.bar[D1_FULL1], 1);
        mbar_init(&s.bar[H_FULL], EPI_WARPS * 32);  mbar_init(&s.bar[H_FULL1], EPI_WARPS * 32);  mbar_init(&s.bar[H_EMPTY], 1);
        for (int q = 0; q < 2; ++q) {
            mbar_init(&s.bar[D2_FULL + q], 1); mbar_init(&s.bar[D2_EMPTY + q], FIN_WARPS * 32); mbar_init(&s.bar[FIN_FULL + q], EPI_WARPS * 32);
        }
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(&s.tmem_base, TMEM_COLS);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = s.tmem_base;
    SPHNCA_TL_PHASE(2);

    if (is_loader) {
        // ===================== loader: rows [A | Y] -> tf32 hi/lo A tile, one tile ahead =====================
        // Lane mapping: a warp owns 32 tile rows; one warp-level LDG.128 fetches 4 consecutive 16-byte chunks of EIGHT rows
        // (lane = 8 * chunk + row): 64 contiguous bytes per row, 8 lines per instruction.  With a thread per row (12 loads of
        // 32 different lines each) the loader alone put 1536 L1 wavefronts per tile on the LSU pipe -- which the epilogue's
        // LDS, the finalizer's stores and the tensor-core issue all share -- against 384 this way.  The shared-memory side is
        // unchanged: the 8 lanes of a quarter warp write 8 consecutive rows of one K chunk, 128 contiguous bytes.
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            if (it > 0) load_rows(tile);
            if (tid == LOAD_WARP0 * 32) SPHNCA_TL(12);
            mbar_wait(&s.bar[Y_EMPTY], (it & 1) ^ 1);
            if (tid == LOAD_WARP0 * 32) SPHNCA_TL(13);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int m = 32 * lw + 8 * g + r8;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float4 x = v[3 * g + q];
                    float hi[4], lo[4];
                    split_tf32(x.x, hi[0], lo[0]); split_tf32(x.y, hi[1], lo[1]);
                    split_tf32(x.z, hi[2], lo[2]); split_tf32(x.w, hi[3], lo[3]);
                    const uint32_t o = chunk_off(TILE, m, (q == 0 ? 0 : 4 * q) + c4) / 4;      // K chunks 0..3 = A, 4..11 = Y
                    *reinterpret_cast<float4*>(&s.Y_hi[o]) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(&s.Y_lo[o]) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(&s.bar[Y_FULL]);
            if (tid == LOAD_WARP0 * 32) SPHNCA_TL(0);
        }
    } else if (warp == MMA_WARP) {
        {   // warp-uniform issue: every lane runs the loop, one elected lane issues (tc_common.cuh)
            // Software pipeline over tiles.  MMAs execute in issue order, so a GEMM1 may be issued into a hidden half's region
            // right behind the GEMM2 MMAs that read it as H_hi:
            //   G2a(i) | G1a(i+1) | G2b(i) | G1b(i+1) | G2a(i+1) ...
            // GEMM1 of the next tile runs while the epilogue warps convert the other half, and each GEMM2 half issues its
            // H_lo MMAs first and releases the shared H_lo region (H_EMPTY) before its H_hi MMAs.
            const uint32_t go = elect_one();
            const uint32_t idesc1 = make_idesc_tf32(TILE, 128), idesc2 = make_idesc_tf32(TILE, N2), idesc2c = make_idesc_tf32(TILE, 2 * N2);
            // base descriptors, built once (desc_advance, tc_common.cuh)
            const uint64_t dYh = make_smem_desc(smem_u32(s.Y_hi), TILE * 16, 128), dYl = make_smem_desc(smem_u32(s.Y_lo), TILE * 16, 128);
            const uint64_t dW1h[2] = {make_smem_desc(smem_u32(s.W1_hi[0]), 128 * 16, 128), make_smem_desc(smem_u32(s.W1_hi[1]), 128 * 16, 128)};
            const uint64_t dW1l[2] = {make_smem_desc(smem_u32(s.W1_lo[0]), 128 * 16, 128), make_smem_desc(smem_u32(s.W1_lo[1]), 128 * 16, 128)};
            const uint64_t dW2[2] = {make_smem_desc(smem_u32(s.W2[0]), 2 * N2 * 16, 128), make_smem_desc(smem_u32(s.W2[1]), 2 * N2 * 16, 128)};
            // The issue loops are ROLLED (a few MMAs per iteration): 136 unrolled MMAs are 7 KB of straight-line code that the
            // MMA warp walks through once per tile, through an instruction cache it shares with four other roles.
            auto gemm1 = [&](const int half) {
                const uint32_t d1 = tbase + (half ? COL_DB : COL_DA);
                const uint64_t wh = dW1h[half], wl = dW1l[half];
#pragma unroll 1
                for (int ks = 0; ks < NIN / 8; ++ks) {
                    const uint32_t oa = chunk_off(TILE, 0, 2) * ks, ob = chunk_off(128, 0, 2) * ks;
                    const uint64_t a_hi = desc_advance(dYh, oa), a_lo = desc_advance(dYl, oa);
                    const uint64_t b_hi = desc_advance(wh, ob), b_lo = desc_advance(wl, ob);
                    mma_tf32_ss_e(d1, a_lo, b_hi, idesc1, ks > 0, go);
                    mma_tf32_ss_e(d1, a_hi, b_lo, idesc1, true, go);
                    mma_tf32_ss_e(d1, a_hi, b_hi, idesc1, true, go);
                }
                mma_commit_e(&s.bar[D1_FULL + half], go);
            };
            auto gemm2 = [&](const int half, const uint32_t d2) {
                const uint32_t hhi = tbase + (half ? COL_DB : COL_DA);
                const uint64_t w = dW2[half];
                // rows 0..31 of a W2 tile = hi, all 64 rows = the stacked [hi ; lo] tile.  H_lo part first; the very first MMA of
                // a tile initialises all 64 columns of the slot (its extra H_lo . W_lo block is the term the split drops anyway)
                if (half == 0) mma_tf32_ts_e(d2, tbase + COL_HLO, w, idesc2c, false, go);
#pragma unroll 1
                for (int ks = (half == 0 ? 1 : 0); ks < 128 / 8; ++ks) {
                    // H_lo . W_hi goes to the SECOND column group (with H_hi . W_lo): the first group then accumulates the large
                    // H_hi . W_hi terms only -- half as many adds in the chain whose TRUNCATING fp32 accumulation biases the result
                    mma_tf32_ts_e(d2 + N2, tbase + COL_HLO + 8 * ks, desc_advance(w, chunk_off(2 * N2, 0, 2) * ks), idesc2, true, go);
                }
                mma_commit_e(&s.bar[H_EMPTY], go);                                                      // H_lo region free again
#pragma unroll 2
                for (int ks = 0; ks < 128 / 8; ++ks)
                    mma_tf32_ts_e(d2, hhi + 8 * ks, desc_advance(w, chunk_off(2 * N2, 0, 2) * ks), idesc2c, true, go);   // H_hi . [W_hi ; W_lo]
            };
            int it = 0;
            if (blockIdx.x < num_tiles) {
                mbar_wait(&s.bar[Y_FULL], 0);
                fence_after_sync();
                gemm1(0); gemm1(1);
                mma_commit_e(&s.bar[Y_EMPTY], go);
            }
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const bool next = tile + (int)gridDim.x < num_tiles;
                const int slot = it & 1;                      // D2 is double buffered: the finalizer warps drain tile it-1 meanwhile
                mbar_wait(&s.bar[D2_EMPTY + slot], ((it >> 1) & 1) ^ 1);
                fence_after_sync();
                const uint32_t d2 = tbase + COL_D2 + 2 * N2 * slot;
                mbar_wait(&s.bar[H_FULL], it & 1);
                fence_after_sync();
                SPHNCA_TL(1);
                gemm2(0, d2);
                SPHNCA_TL(4);
                if (next) {
                    mbar_wait(&s.bar[Y_FULL], (it + 1) & 1);
                    fence_after_sync();
                    gemm1(0);
                    SPHNCA_TL(2);
                }
                mbar_wait(&s.bar[H_FULL1], it & 1);
                fence_after_sync();
                gemm2(1, d2);
                mma_commit_e(&s.bar[D2_FULL + slot], go);
                SPHNCA_TL(5);
                if (next) {
                    gemm1(1);
                    mma_commit_e(&s.bar[Y_EMPTY], go);
                    SPHNCA_TL(3);
                }
            }
        }
    } else if (warp >= FIN_WARP0) {
        // ===================== finalizer: 4 warps, thread = tile row; gates + update + the row stores =====================
        // Its own warps, so that the update of tile i (~3 k cycles, bound by the 17 row-strided 16-byte stores per row) runs
        // beside the MMAs and the hidden-layer epilogue of tile i+1 instead of in front of them (D2 has two slots).
        const int m = (warp & 3) * 32 + lane;                 // a warp may only touch TMEM lanes 32*(warp % 4)..+31
        const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int j = tile * TILE + m;
            const bool valid = j < n_total;
            const int p = (valid && rows) ? __ldg(rows + j) : j;
            const int slot = it & 1;
            // fetched before the waits (raw asm loads: nothing consumes them until the update below)
            float4 Ai4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) Ai4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t fire_raw = 0, flag_raw = 0;
            if (valid) {
                const float4* a4 = reinterpret_cast<const float4*>(A_s + (size_t)(in_by_list ? j : p) * NF);
#pragma unroll
                for (int k = 0; k < 4; ++k) Ai4[k] = ldg_f4_raw(a4 + k);
                fire_raw = rows ? 1u : ldg_u8_raw(fire_s + p);
                flag_raw = ldg_u8_raw(flags_s + p);
            }
            mbar_wait(&s.bar[FIN_FULL + slot], (it >> 1) & 1);
            const float o32sum = s.o32[slot][0][m] + s.o32[slot][1][m];      // fixed order: group 0 + group 1
            mbar_wait(&s.bar[D2_FULL + slot], (it >> 1) & 1);
            fence_after_sync();
            if (tid == (FIN_WARP0 + 3) * 32) SPHNCA_TL(10);
            const uint32_t d2 = lane_addr + COL_D2 + 2 * N2 * slot;
            float out[NO_PAD];
            {
                uint32_t r[16], r2[16];                       // hi.(hi+lo of H) group and hi.lo group of the stacked GEMM2
                tmem_ld16(d2, r); tmem_ld16(d2 + N2, r2); wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) out[j] = (__uint_as_float(r[j]) + __uint_as_float(r2[j])) + s.b2[j];
                tmem_ld16(d2 + 16, r); tmem_ld16(d2 + N2 + 16, r2); wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) out[16 + j] = (__uint_as_float(r[j]) + __uint_as_float(r2[j])) + s.b2[16 + j];
            }
            fence_before_sync();
            mbar_arrive(&s.bar[D2_EMPTY + slot]);
            if (tid == (FIN_WARP0 + 3) * 32) SPHNCA_TL(14);
            out[32] = o32sum + s.b2[32]; out[33] = 0.f; out[34] = 0.f; out[35] = 0.f;
            if (c.rule != 0) {
#pragma unroll
                for (int o = NF; o < NO_PAD; ++o) out[o] = 0.f;
            }
            if (valid) {
                float Ai[NF], nA[NF];
#pragma unroll
                for (int k = 0; k < 4; ++k) { Ai[4 * k] = Ai4[k].x; Ai[4 * k + 1] = Ai4[k].y; Ai[4 * k + 2] = Ai4[k].z; Ai[4 * k + 3] = Ai4[k].w; }
                const bool fire = fire_raw != 0;
                if (c.rule == 0) {
                    const float mult = fast_sigmoid(out[2 * NF]);
#pragma unroll
                    for (int f = 0; f < NF; ++f) nA[f] = fmaf(Ai[f], fast_sigmoid(out[f]), fast_tanh(out[NF + f]) * mult);
                } else {
#pragma unroll
                    for (int f = 0; f < NF; ++f) nA[f] = fmaf(out[f], c.orig_rate, Ai[f]);
                }
#pragma unroll
                for (int f = 0; f < NF; ++f) nA[f] = fire ? nA[f] : Ai[f];
                if (tid == (FIN_WARP0 + 3) * 32) SPHNCA_TL(15);
                if (fold_mask && !(flag_raw & 1u)) {
#pragma unroll
                    for (int f = 0; f < NF; ++f) nA[f] *= 0.f;
                }
                float4* o4 = reinterpret_cast<float4*>(An_s + (size_t)p * NF);
#pragma unroll
                for (int k = 0; k < 4; ++k) o4[k] = make_float4(nA[4 * k], nA[4 * k + 1], nA[4 * k + 2], nA[4 * k + 3]);
                // fold_mask (no alpha channel, sphnca_fire_list_draw set bits 0|1 of the living rows): the static life mask is applied
                // here instead of by step_mask_static_kernel -- dead rows are multiplied by 0 exactly like there
                flags_s[p] = (uint8_t)((flag_raw & (fold_mask ? 3u : 1u)) | (fire ? 4u : 0u));
                alive_s[p] = c.use_alpha ? (nA[3] > 0.1f ? 1 : 0) : 1;
            }
            // The pre-gate outputs O (36 floats per row, rows consecutive in j) leave through shared memory: 16 rows of the warp
            // at a time are laid down as they are in global memory (row pitch 144 B: the 8 lanes of a quarter warp hit distinct
            // banks) and go out as 512 contiguous bytes per warp-level STG.128 -- 36 L1 wavefronts and 36 full-line L2 writes per
            // warp and tile instead of 288 wavefronts and 288 half-sector writes with a thread per row.
            {
                float4* st4 = reinterpret_cast<float4*>(s.ostage[warp - FIN_WARP0]);
                const int jw = tile * TILE + (warp & 3) * 32;
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    if ((lane >> 4) == sub) {
#pragma unroll
                        for (int k = 0; k < NO_PAD / 4; ++k) st4[(lane & 15) * (NO_PAD / 4) + k] = make_float4(out[4 * k], out[4 * k + 1], out[4 * k + 2], out[4 * k + 3]);
                    }
                    __syncwarp();
                    const int j0 = jw + sub * 16;
                    float4* g4 = reinterpret_cast<float4*>(O_s + (size_t)j0 * NO_PAD);
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const int cidx = k * 32 + lane;
                        if (cidx < 16 * (NO_PAD / 4) && j0 + cidx / (NO_PAD / 4) < n_total) g4[cidx] = st4[cidx];
                    }
                    __syncwarp();
                }
            }
            if (tid == (FIN_WARP0 + 3) * 32) SPHNCA_TL(11);
        }
    } else {
        // ===================== epilogue: 8 warps; warp & 3 = TMEM lane quarter, warp >> 2 = column half =====================
        // group chalf converts hidden columns [64*chalf, 64*chalf+64) of each 128 half: bias + ReLU, the ReLU bits, the partial
        // scalar-gate dot product, H as tf32 hi/lo back to TMEM (the A operand of GEMM2)
        const int m = (warp & 3) * 32 + lane;                 // tile row == TMEM lane
        const int chalf = warp >> 2;
        const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int j = tile * TILE + m;
            const bool valid = j < n_total;
            float2 o32p = make_float2(0.f, 0.f);              // even / odd hidden units of the scalar-gate dot product
            uint32_t rmask[NH / 64];                          // this warp's 4 ReLU words: (half, chalf) -> word 4*half + 2*chalf + {0,1}
#pragma unroll
            for (int wd = 0; wd < NH / 64; ++wd) rmask[wd] = 0u;
            // One 16-column block: bias + ReLU (+ its bits), the partial scalar-gate dot product, H as tf32 hi / lo.
            // Packed fp32 pairs (FADD2 / FFMA2) for the bias add, the dot product and the lo remainder.
            auto convert = [&](const uint32_t (&r)[16], const float* __restrict__ b1p, const float* __restrict__ w2p,
                               uint32_t (&hi)[16], uint32_t (&lo)[16]) -> uint32_t {
                uint32_t bits = 0;
                // descending j: the ReLU bit of element j is the SIGN of (0 - pre) -- set exactly when pre > 0, for +-0 too -- shifted
                // into the word by one funnel shift per element (element 0 last: bit 0); bias / gate weights come four at a time
#pragma unroll
                for (int j4 = 12; j4 >= 0; j4 -= 4) {
                    const float4 bq = *reinterpret_cast<const float4*>(b1p + j4), wq = *reinterpret_cast<const float4*>(w2p + j4);
#pragma unroll
                    for (int jj = 2; jj >= 0; jj -= 2) {
                        const int j = j4 + jj;
                        const float2 pre = fadd2(make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), jj ? make_float2(bq.z, bq.w) : make_float2(bq.x, bq.y));
                        const float2 npre = fsub2(make_float2(0.f, 0.f), pre);
                        bits = __funnelshift_l(__float_as_uint(npre.y), bits, 1);
                        bits = __funnelshift_l(__float_as_uint(npre.x), bits, 1);
                        const float2 hv = make_float2(fmaxf(pre.x, 0.f), fmaxf(pre.y, 0.f));
                        o32p = ffma2(hv, jj ? make_float2(wq.z, wq.w) : make_float2(wq.x, wq.y), o32p);
                        const float2 h_ = make_float2(__uint_as_float((__float_as_uint(hv.x) + 0x1000u) & 0xffffe000u),
                                                      __uint_as_float((__float_as_uint(hv.y) + 0x1000u) & 0xffffe000u));   // split_tf32
                        const float2 l_ = fsub2(hv, h_);
                        hi[j] = __float_as_uint(h_.x); hi[j + 1] = __float_as_uint(h_.y);
                        lo[j] = __float_as_uint(l_.x); lo[j + 1] = __float_as_uint(l_.y);
                    }
                }
                return bits;
            };
            // ROLLED loops over the two hidden halves and the two 32-column pairs of a half: the conversion code exists twice
            // (one copy per register set of the TMEM-load double buffer), not eight times.  Fully unrolled, the epilogue alone
            // was 18 KB of straight-line code per tile, five roles share each scheduler's instruction cache, and ncu showed
            // "no instruction" as the epilogue warps' first stall reason (25 % of their samples).
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const uint32_t d1 = lane_addr + COL_DA + half * (COL_DB - COL_DA);
                mbar_wait(&s.bar[D1_FULL + half], it & 1);
                fence_after_sync();
                if (tid == 0) SPHNCA_TL(6 + 2 * half);
                uint32_t r0[16], r1[16];                      // the next 16 columns are in flight while the current ones are converted
                tmem_ld16(d1 + chalf * 64, r0);
                uint32_t word[2];
#pragma unroll 1
                for (int pair = 0; pair < 2; ++pair) {
                    const int cc = chalf * 64 + pair * 32;
                    const float* b1p = s.b1 + half * 128 + cc;
                    const float* w2p = s.w2_last + half * 128 + cc;
                    uint32_t hi[16], lo[16];
                    wait_ld();
                    tmem_ld16(d1 + cc + 16, r1);
                    const uint32_t bits_a = convert(r0, b1p, w2p, hi, lo);
                    tmem_st16(d1 + cc, hi);                   // in place: pre-activations -> H_hi
                    if (pair == 0) {                          // the shared H_lo region: free once the previous GEMM2 half's H_lo MMAs are done
                        mbar_wait(&s.bar[H_EMPTY], half ^ 1);
                        fence_after_sync();
                    }
                    tmem_st16(lane_addr + COL_HLO + cc, lo);
                    wait_ld();
                    if (pair == 0) tmem_ld16(d1 + cc + 32, r0);
                    const uint32_t bits_b = convert(r1, b1p + 16, w2p + 16, hi, lo);
                    tmem_st16(d1 + cc + 16, hi);
                    tmem_st16(lane_addr + COL_HLO + cc + 16, lo);
                    const uint32_t w = bits_a | (bits_b << 16);
                    if (pair == 0) word[0] = w; else word[1] = w;
                }
                if (half == 0) { rmask[0] = word[0]; rmask[1] = word[1]; } else { rmask[2] = word[0]; rmask[3] = word[1]; }
                wait_st();
                fence_before_sync();
                mbar_arrive(&s.bar[H_FULL + half]);
                if (tid == 0) SPHNCA_TL(7 + 2 * half);
            }
            // hand the partial scalar-gate sum to the row's finalizer thread.  The slot was last read by the finalizer of tile
            // it-2, before it released D2's slot; GEMM2 of this tile waited for that release and this epilogue for GEMM2.
            const int slot = it & 1;
            s.o32[slot][chalf][m] = o32p.x + o32p.y;
            mbar_arrive(&s.bar[FIN_FULL + slot]);
            if (valid) {
                // ReLU words: hidden 0..63 -> words 0,1 (group 0); 64..127 -> 2,3 (group 1); 128..191 -> 4,5 (group 0); 192..255 -> 6,7 (group 1)
                uint2* rm2 = reinterpret_cast<uint2*>(relu_s + (size_t)j * (NH / 32));
                rm2[chalf] = make_uint2(rmask[0], rmask[1]);
                rm2[2 + chalf] = make_uint2(rmask[2], rmask[3]);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    SPHNCA_TL_PHASE(3);
    if (warp == MMA_WARP) tmem_dealloc(tbase, mt::TMEM_COLS);
}

int launch_step_mlp_tc(int64_t N, const StepConsts& c, const sphnca_step_params* p, const float* A_s, const float* Y_s,
                       const uint8_t* fire_s, float* A_new_s, uint8_t* alive_s, float* O_s, uint8_t* flags_s,
                       uint32_t* relu_s, cudaStream_t stream, const int32_t* rows, const int32_t* count, float* A_c, float* Y_c,
                       int in_by_list, int fold_mask) {
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        SPHNCA_CUDA(cudaGetDevice(&dev));
        SPHNCA_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int num_tiles = (int)((N + mt::TILE - 1) / mt::TILE);
    const int blocks = num_tiles < num_sms ? num_tiles : num_sms;
    const size_t smem = sizeof(mt::Smem);
    SPHNCA_CUDA(cudaFuncSetAttribute(step_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SPHNCA_LAUNCH(step_mlp_tc_kernel, blocks, mt::THREADS, smem, stream, (int)N, A_s, Y_s, fire_s, *p, c, A_new_s, O_s, flags_s,
                  alive_s, relu_s, num_tiles, rows, count, A_c, Y_c, in_by_list, fold_mask);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

}  // namespace sphnca

SPHNCA_TIMELINE_READER(sphnca_debug_timeline_mlp)
