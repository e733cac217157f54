// Data gradient of the NCA step after the perception sweep, on the Blackwell tensor path.
// Same maths as step_bwd_kernel (step.cu); no hidden-layer recompute: the forward saved the ReLU
// activity bits (32 B / particle).
//
//   loader warps (4)   thread = particle: gate / tanh / fire / life-mask derivatives -> dO[36] (written to
//                      HBM for the weight-gradient stage and, tf32 hi/lo split, into the MMA A tile),
//                      the direct part of dL/dA, the scalar-gate column dO[32], the ReLU bits -> smem.
//   MMA warp (1 lane)  GEMM3  D3[128p x 128h] = dO[128 x 32] . W2t_half^T      (B = W2^T stored K-major)
//                      GEMM4  Dy[128p x 48]  += dH[128 x 128](TMEM) . W1t_half^T (B = W1^T stored K-major)
//   epilogue warps (8) tcgen05.ld D3 -> (+ dO[32]*W2[32][h]) -> ReLU bits -> tf32 hi/lo -> tcgen05.st dH
//                      IN PLACE over D3 (hi) and into the shared dH_lo region.
//   finalizer warps (4) tcgen05.ld Dy -> dL/dY (32) and dL/dA direct part (16): the row stores of tile i run
//                      beside the MMAs and the dH epilogue of tile i+1.
//   The MMA warp runs a software pipeline over tiles (MMAs execute in issue order):
//       G4a(i) | G3a(i+1) | G4b(i) | G3b(i+1) | G4a(i+1) ...
//   each hidden half has its own D3/dH_hi region, GEMM3 of the next tile is issued right behind the GEMM4 MMAs that
//   read the region, and each GEMM4 half issues its dH_lo MMAs first and then releases the dH_lo region.
//
// (MN-major tf32 operands need the SW128_32B layout, so instead of re-viewing the forward's W1/W2 tiles
//  the kernel stages transposed K-major copies: W2^T 64 KB + W1^T 96 KB, hi/lo.)
// GEMM4 has only 48 output columns, where an MMA costs ~32 cycles + N/4 whatever N is (probes/tc_timing4: A from
// TMEM, N=48 59 cycles, N=128 71), so the hi and lo copies of W1^T are STACKED along N (96 rows): dH_hi . [W_hi ; W_lo]
// is one MMA into Dy columns 0..47 | 48..95 and dH_lo . W_hi adds to columns 0..47 -- 2 MMAs per k-step instead of 3;
// the final epilogue adds the two column groups.
// TMEM: D3/dH_hi 2 x 128 + dH_lo 128 + Dy 96 = 480 columns.
#include <type_traits>
#include "step_common.cuh"
#include "tc_common.cuh"

namespace sphnca {

using namespace tc;

namespace bw {
constexpr int TILE = 128;
// epilogue warps 0..7: warp & 3 = TMEM lane quarter, warp >> 2 = which 64 of a half's 128 hidden columns
constexpr int EPI_WARPS = 8, MMA_WARP = 8, LOAD_WARP0 = 9, LOAD_WARPS = 4, FIN_WARP0 = 13, FIN_WARPS = 4;
constexpr int THREADS = (FIN_WARP0 + FIN_WARPS) * 32;        // 544
constexpr int KO = 32;                                       // outputs handled by the tensor core
constexpr uint32_t COL_DA = 0, COL_DB = 128, COL_DHLO = 256, COL_DY = 384, TMEM_COLS = 512;
enum { DO_FULL = 0, DO_EMPTY, D3_FULL, D3_FULL1, DH_FULL, DH_FULL1, DH_EMPTY, DY_FULL, DY_EMPTY, NUM_BARS };

struct Smem {
    float W2t_hi[2][128 * KO], W2t_lo[2][128 * KO];       // [hidden half] rows = h (128), K = o (32)
    float W1t[2][2 * NIN * 128];                          // [hidden half] rows 0..47 = k hi, 48..95 = k lo ; K = h (128)
    float dO_hi[TILE * KO], dO_lo[TILE * KO];             // rows = particle, K = o
    float dAdir[2][TILE][NF];
    float dO32[2][TILE];
    uint32_t relu[2][TILE][NH / 32];
    float w2_last[NH];
    float fstage[FIN_WARPS][8 * 52];                      // finalizer: 8 result rows [dAdir 16 | dY 32 | particle, pad 3] per warp, on their way out
    uint64_t bar[NUM_BARS];
    uint32_t tmem_base;
};
__device__ __forceinline__ uint32_t fidx(uint32_t rows, uint32_t m, uint32_t k) { return chunk_off(rows, m, k >> 2) / 4 + (k & 3); }
}  // namespace bw

__global__ void __launch_bounds__(bw::THREADS, 1)
step_bwd_tc_kernel(int n_total, const float* __restrict__ A_s, const float* __restrict__ O_s,
                   const uint8_t* __restrict__ flags_s, const uint32_t* __restrict__ relu_s,
                   const float* __restrict__ dAn_s, sphnca_step_params prm, StepConsts c,
                   float* __restrict__ dO_s, float* __restrict__ dY_s, float* __restrict__ dAdir_s, int num_tiles,
                   const int32_t* __restrict__ rows, const int32_t* __restrict__ count_ptr) {
    using namespace bw;
    // Fire-list mode (rows != NULL, firelist.cu): tile row j stands for sorted particle rows[j], j < *count.  A_s (the compact
    // copy A_c), O_s, relu_s and dO_s are indexed by j; flags, dL/dA_new, dL/dY and the direct dL/dA part by the particle.
    if (count_ptr != nullptr) {
        n_total = __ldg(count_ptr);
        num_tiles = (n_total + TILE - 1) / TILE;
    }
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int n_mma = c.n_out < KO ? c.n_out : KO;
    // W2 [o][h] -> W2t[half] (rows h, K = o)
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        // W2 [o][h] -> W2t_hi / W2t_lo [half]: rows = h, K = o (4 global rows per chunk, consecutive lanes along h)
        stage_chunks<2>(128 * (KO / 4), tid, THREADS, [&](int idx) {
            int row, ck; stage_map(idx, 128 / 8, row, ck);
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = 4 * ck + e < n_mma ? __ldg(prm.W2 + (size_t)(4 * ck + e) * NH + half * 128 + row) : 0.f;
            return make_float4(v[0], v[1], v[2], v[3]);
        }, [&](int idx, float4 hi, float4 lo) {
            int row, ck; stage_map(idx, 128 / 8, row, ck);
            const uint32_t o = chunk_off(128, row, ck) / 4;
            *reinterpret_cast<float4*>(&s.W2t_hi[half][o]) = hi; *reinterpret_cast<float4*>(&s.W2t_lo[half][o]) = lo;
        });
        // W1 [h][k] -> W1t[half]: rows 0..47 = k hi, 48..95 = k lo, K = h within the half
        stage_chunks<3>(NIN * (128 / 4), tid, THREADS, [&](int idx) {
            int row, ck; stage_map(idx, NIN / 8, row, ck);
            const float* w = prm.W1 + (size_t)(half * 128 + 4 * ck) * NIN + row;
            return make_float4(__ldg(w), __ldg(w + NIN), __ldg(w + 2 * NIN), __ldg(w + 3 * NIN));
        }, [&](int idx, float4 hi, float4 lo) {
            int row, ck; stage_map(idx, NIN / 8, row, ck);
            *reinterpret_cast<float4*>(&s.W1t[half][chunk_off(2 * NIN, row, ck) / 4]) = hi;
            *reinterpret_cast<float4*>(&s.W1t[half][chunk_off(2 * NIN, NIN + row, ck) / 4]) = lo;
        });
    }
    for (int i = tid; i < NH; i += THREADS) s.w2_last[i] = c.n_out > KO ? __ldg(prm.W2 + (size_t)KO * NH + i) : 0.f;
    if (tid == 0) {
        mbar_init(&s.bar[DO_FULL], LOAD_WARPS * 32); mbar_init(&s.bar[DO_EMPTY], 1);
        mbar_init(&s.bar[D3_FULL], 1);               mbar_init(&s.bar[D3_FULL1], 1);
        mbar_init(&s.bar[DH_FULL], EPI_WARPS * 32);  mbar_init(&s.bar[DH_FULL1], EPI_WARPS * 32);  mbar_init(&s.bar[DH_EMPTY], 1);
        mbar_init(&s.bar[DY_FULL], 1);               mbar_init(&s.bar[DY_EMPTY], FIN_WARPS * 32);
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(&s.tmem_base, TMEM_COLS);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = s.tmem_base;

    if (warp >= LOAD_WARP0 && warp < FIN_WARP0) {
        // ======================= prologue: dL/d(MLP output) and the direct dL/dA part =======================
        const int m = tid - LOAD_WARP0 * 32;
        int it = 0;
        // fire-list mode: the row's particle index is fetched ONE TILE AHEAD -- the loads below depend on it, and a cold list
        // read in front of them put a second global-memory round trip on this role's critical path (+8 us per 262,144 rows)
        int p_next = (int)blockIdx.x * TILE + m;
        if (rows && p_next < n_total) p_next = __ldg(rows + p_next);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int j = tile * TILE + m;
            const bool valid = j < n_total;
            const int p = p_next;
            {
                const int jn = j + (int)gridDim.x * TILE;
                p_next = (rows && jn < n_total) ? __ldg(rows + jn) : jn;
            }
            float dO[NO_PAD], dAdir[NF];
            uint32_t rbits[NH / 32];
#pragma unroll
            for (int o = 0; o < NO_PAD; ++o) dO[o] = 0.f;
#pragma unroll
            for (int f = 0; f < NF; ++f) dAdir[f] = 0.f;
#pragma unroll
            for (int w = 0; w < NH / 32; ++w) rbits[w] = 0u;
            if (valid) {
                float Ai[NF], dnA[NF];
                load16(A_s + (size_t)j * NF, Ai);
                load16(dAn_s + (size_t)p * NF, dnA);
                const uint8_t fl = flags_s[p];
                const bool living = (fl & 1) && (fl & 2), fire = fl & 4;
#pragma unroll
                for (int f = 0; f < NF; ++f) dnA[f] = living ? dnA[f] : 0.f;
                if (fire) {
                    const float4* o4 = reinterpret_cast<const float4*>(O_s + (size_t)j * NO_PAD);
                    float out[NO_PAD];
#pragma unroll
                    for (int k = 0; k < NO_PAD / 4; ++k) { float4 t = __ldg(o4 + k); out[4*k] = t.x; out[4*k+1] = t.y; out[4*k+2] = t.z; out[4*k+3] = t.w; }
                    if (c.rule == 0) {
                        // fast exponential (2 ulp): |error| ~1e-7 on the gates, far inside the parity bar
                        const float mult = __fdividef(1.f, 1.f + __expf(-out[2 * NF]));
                        float dmult = 0.f;
#pragma unroll
                        for (int f = 0; f < NF; ++f) {
                            const float gate = __fdividef(1.f, 1.f + __expf(-out[f]));
                            const float e2 = __expf(-2.f * fabsf(out[NF + f]));
                            const float delta = copysignf(__fdividef(1.f - e2, 1.f + e2), out[NF + f]);
                            dO[f] = dnA[f] * Ai[f] * gate * (1.f - gate);
                            dO[NF + f] = dnA[f] * mult * (1.f - delta * delta);
                            dmult = fmaf(dnA[f], delta, dmult);
                            dAdir[f] = dnA[f] * gate;
                        }
                        dO[2 * NF] = dmult * mult * (1.f - mult);
                    } else {
#pragma unroll
                        for (int f = 0; f < NF; ++f) { dO[f] = dnA[f] * c.orig_rate; dAdir[f] = dnA[f]; }
                    }
                    const uint4* r4 = reinterpret_cast<const uint4*>(relu_s + (size_t)j * (NH / 32));
                    const uint4 r0 = __ldg(r4), r1 = __ldg(r4 + 1);
                    rbits[0] = r0.x; rbits[1] = r0.y; rbits[2] = r0.z; rbits[3] = r0.w;
                    rbits[4] = r1.x; rbits[5] = r1.y; rbits[6] = r1.z; rbits[7] = r1.w;
                } else {
#pragma unroll
                    for (int f = 0; f < NF; ++f) dAdir[f] = dnA[f];
                }
                float4* do4 = reinterpret_cast<float4*>(dO_s + (size_t)j * NO_PAD);
#pragma unroll
                for (int k = 0; k < NO_PAD / 4; ++k) do4[k] = make_float4(dO[4*k], dO[4*k+1], dO[4*k+2], dO[4*k+3]);
            }
            // GEMM3 of the previous tile consumed the dO tile and the finalizer of tile it-2 this slot's dAdir rows
            mbar_wait(&s.bar[DO_EMPTY], (it & 1) ^ 1);
#pragma unroll
            for (int ch = 0; ch < KO / 4; ++ch) {
                float hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) split_tf32(dO[4 * ch + j], hi[j], lo[j]);
                const uint32_t o = chunk_off(TILE, m, ch) / 4;
                *reinterpret_cast<float4*>(&s.dO_hi[o]) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(&s.dO_lo[o]) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
            const int slot = it & 1;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                *reinterpret_cast<float4*>(&s.dAdir[slot][m][4 * k]) = make_float4(dAdir[4*k], dAdir[4*k+1], dAdir[4*k+2], dAdir[4*k+3]);
            s.dO32[slot][m] = dO[KO];
            *reinterpret_cast<uint4*>(&s.relu[slot][m][0]) = make_uint4(rbits[0], rbits[1], rbits[2], rbits[3]);
            *reinterpret_cast<uint4*>(&s.relu[slot][m][4]) = make_uint4(rbits[4], rbits[5], rbits[6], rbits[7]);
            fence_proxy_async_smem();
            mbar_arrive(&s.bar[DO_FULL]);
        }
    } else if (warp == MMA_WARP) {
        {   // warp-uniform issue: every lane runs the loop, one elected lane issues (tc_common.cuh)
            const uint32_t go = elect_one();
            const uint32_t idesc3 = make_idesc_tf32(TILE, 128), idesc4 = make_idesc_tf32(TILE, NIN), idesc4c = make_idesc_tf32(TILE, 2 * NIN);
            // base descriptors, built once (desc_advance, tc_common.cuh)
            const uint64_t dOh = make_smem_desc(smem_u32(s.dO_hi), TILE * 16, 128), dOl = make_smem_desc(smem_u32(s.dO_lo), TILE * 16, 128);
            const uint64_t dW2h[2] = {make_smem_desc(smem_u32(s.W2t_hi[0]), 128 * 16, 128), make_smem_desc(smem_u32(s.W2t_hi[1]), 128 * 16, 128)};
            const uint64_t dW2l[2] = {make_smem_desc(smem_u32(s.W2t_lo[0]), 128 * 16, 128), make_smem_desc(smem_u32(s.W2t_lo[1]), 128 * 16, 128)};
            const uint64_t dW1t[2] = {make_smem_desc(smem_u32(s.W1t[0]), 2 * NIN * 16, 128), make_smem_desc(smem_u32(s.W1t[1]), 2 * NIN * 16, 128)};
            auto gemm3 = [&](const int half) {
                const uint32_t d3 = tbase + (half ? COL_DB : COL_DA);
#pragma unroll
                for (int ks = 0; ks < KO / 8; ++ks) {
                    const uint64_t a_hi = desc_advance(dOh, chunk_off(TILE, 0, 2 * ks)), a_lo = desc_advance(dOl, chunk_off(TILE, 0, 2 * ks));
                    const uint64_t b_hi = desc_advance(dW2h[half], chunk_off(128, 0, 2 * ks)), b_lo = desc_advance(dW2l[half], chunk_off(128, 0, 2 * ks));
                    mma_tf32_ss_e(d3, a_lo, b_hi, idesc3, ks > 0, go);
                    mma_tf32_ss_e(d3, a_hi, b_lo, idesc3, true, go);
                    mma_tf32_ss_e(d3, a_hi, b_hi, idesc3, true, go);
                }
                mma_commit_e(&s.bar[D3_FULL + half], go);
            };
            auto gemm4 = [&](const int half) {
                const uint32_t dhhi = tbase + (half ? COL_DB : COL_DA);
                // rows 0..47 of a W1^T tile = hi, all 96 rows = the stacked [hi ; lo] tile.  dH_lo part first; the very first
                // MMA of a tile initialises all 96 Dy columns (its extra dH_lo . W_lo block is the term the split drops anyway)
#pragma unroll
                for (int ks = 0; ks < 128 / 8; ++ks) {
                    const uint64_t b = desc_advance(dW1t[half], chunk_off(2 * NIN, 0, 2 * ks));
                    if (half == 0 && ks == 0) mma_tf32_ts_e(tbase + COL_DY, tbase + COL_DHLO, b, idesc4c, false, go);
                    else mma_tf32_ts_e(tbase + COL_DY + NIN, tbase + COL_DHLO + 8 * ks, b, idesc4, true, go);   // dH_lo . W_hi -> second group (see step_mlp_tc.cu)
                }
                mma_commit_e(&s.bar[DH_EMPTY], go);                                                             // dH_lo region free again
#pragma unroll
                for (int ks = 0; ks < 128 / 8; ++ks) {
                    const uint64_t b = desc_advance(dW1t[half], chunk_off(2 * NIN, 0, 2 * ks));
                    mma_tf32_ts_e(tbase + COL_DY, dhhi + 8 * ks, b, idesc4c, true, go);                         // dH_hi . [W_hi ; W_lo]
                }
            };
            int it = 0;
            if (blockIdx.x < num_tiles) {
                mbar_wait(&s.bar[DO_FULL], 0);
                fence_after_sync();
                gemm3(0); gemm3(1);
            }
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const bool next = tile + (int)gridDim.x < num_tiles;
                mbar_wait(&s.bar[DH_FULL], it & 1);
                fence_after_sync();
                mbar_wait(&s.bar[DY_EMPTY], (it & 1) ^ 1);    // the finalizer drained Dy (and its dAdir rows) of the previous tile
                fence_after_sync();
                // loader release: GEMM3 of THIS tile (issued one iteration ago) has read the dO tile, and -- by the wait above --
                // the finalizer of tile it-1 is done with its smem slot, the one the loader of tile it+1 writes
                mma_commit_e(&s.bar[DO_EMPTY], go);
                gemm4(0);
                if (next) {
                    mbar_wait(&s.bar[DO_FULL], (it + 1) & 1);
                    fence_after_sync();
                    gemm3(0);
                }
                mbar_wait(&s.bar[DH_FULL1], it & 1);
                fence_after_sync();
                gemm4(1);
                mma_commit_e(&s.bar[DY_FULL], go);
                if (next) gemm3(1);
            }
        }
    } else if (warp >= FIN_WARP0) {
        // ======================= finalizer: thread = tile row; Dy -> dL/dY and the direct dL/dA part =======================
        const int m = (warp & 3) * 32 + lane;                 // a warp may only touch TMEM lanes 32*(warp % 4)..+31
        const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int j = tile * TILE + m;
            const bool valid = j < n_total;
            const int p = (valid && rows) ? __ldg(rows + j) : j;
            const int slot = it & 1;
            mbar_wait(&s.bar[DY_FULL], it & 1);
            fence_after_sync();
            float dy[NIN];
#pragma unroll
            for (int cc = 0; cc < NIN; cc += 16) {
                uint32_t r[16], r2[16];                       // hi.(hi+lo of dH) group and hi.lo group of the stacked GEMM4
                tmem_ld16(lane_addr + COL_DY + cc, r); tmem_ld16(lane_addr + COL_DY + NIN + cc, r2); wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) dy[cc + j] = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
            }
            float4 dir[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) dir[k] = *reinterpret_cast<const float4*>(&s.dAdir[slot][m][4 * k]);
            fence_before_sync();
            mbar_arrive(&s.bar[DY_EMPTY]);                    // Dy and this slot's dAdir rows are in registers
            // The result rows (64 B of dAdir and 128 B of dL/dY at the particle's place) leave through shared memory, 8 rows of the
            // warp at a time (row pitch 208 B: the 8 writing lanes hit distinct banks), as contiguous 16-byte pieces: a warp-level
            // store then touches ~6 lines instead of 32.  With a thread per row these 12 stores cost the kernel 14 of its 78 us
            // (measured by compiling them out) -- the LSU pipe they clog is shared with the prologue's loads and the epilogue's LDS.
            {
                float4* st4 = reinterpret_cast<float4*>(s.fstage[warp - FIN_WARP0]);
                float4* dA4 = reinterpret_cast<float4*>(dAdir_s);
                float4* dY4 = reinterpret_cast<float4*>(dY_s);
                const int jw = tile * TILE + (warp & 3) * 32;
                const bool ilv = rows != nullptr && c.dy_ilv;     // list mode only: sphnca_step_bwd keeps the plain layout
#pragma unroll 1
                for (int sub = 0; sub < 4; ++sub) {
                    if ((lane >> 3) == sub) {
                        float4* r = st4 + (lane & 7) * 13;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            r[k] = make_float4(dir[k].x + dy[4*k], dir[k].y + dy[4*k+1], dir[k].z + dy[4*k+2], dir[k].w + dy[4*k+3]);
#pragma unroll
                        for (int k = 0; k < 8; ++k)      // dy_ilv: [x quarter | y quarter] pairs (sphnca_step_params.dy_interleaved)
                            r[4 + (ilv ? ((k & 3) * 2 + (k >> 2)) : k)] = make_float4(dy[NF+4*k], dy[NF+4*k+1], dy[NF+4*k+2], dy[NF+4*k+3]);
                        reinterpret_cast<int*>(r + 12)[0] = p;            // the row's particle, in the pad piece: no list lookups below
                    }
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int idx = k * 32 + lane, row = idx / 12, cq = idx - 12 * row;
                        const int jr = jw + sub * 8 + row;
                        if (jr < n_total) {
                            const size_t pr = (size_t)reinterpret_cast<const int*>(st4 + row * 13 + 12)[0];
                            const float4 v = st4[row * 13 + cq];
                            if (cq < 4) dA4[pr * 4 + cq] = v; else dY4[pr * 8 + (cq - 4)] = v;
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ======================= epilogue: dH = (D3 + dO[32] W2[32][h]) * ReLU bits, tf32 hi (in place) / lo =======================
        const int m = (warp & 3) * 32 + lane;
        const int chalf = warp >> 2;
        const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int slot = it & 1;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const uint32_t d3 = lane_addr + (half ? COL_DB : COL_DA);
                mbar_wait(&s.bar[D3_FULL + half], it & 1);
                fence_after_sync();
                const float d32 = s.dO32[slot][m];
                // the next 16 columns are in flight while the current ones are converted (two register sets)
                uint32_t r[2][16];
                tmem_ld16(d3 + chalf * 64, r[0]);
                auto block = [&](auto C4) {
                    constexpr int c4 = decltype(C4)::value;
                    const int cc = chalf * 64 + c4 * 16;
                    uint32_t hi[16], lo[16];
                    wait_ld();
                    if (c4 < 3) tmem_ld16(d3 + cc + 16, r[(c4 + 1) & 1]);
                    const uint32_t bits = s.relu[slot][m][(half * 128 + cc) >> 5] >> (cc & 16);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = half * 128 + cc + j;
                        const float dh = ((bits >> j) & 1u) ? fmaf(d32, s.w2_last[n], __uint_as_float(r[c4 & 1][j])) : 0.f;
                        float h_, l_; split_tf32(dh, h_, l_);
                        hi[j] = __float_as_uint(h_); lo[j] = __float_as_uint(l_);
                    }
                    tmem_st16(d3 + cc, hi);                   // in place: D3 -> dH_hi
                    if (c4 == 0) {                            // the shared dH_lo region: free once the previous GEMM4 half's dH_lo MMAs are done
                        mbar_wait(&s.bar[DH_EMPTY], half ^ 1);
                        fence_after_sync();
                    }
                    tmem_st16(lane_addr + COL_DHLO + cc, lo);
                };
                block(std::integral_constant<int, 0>{}); block(std::integral_constant<int, 1>{});
                block(std::integral_constant<int, 2>{}); block(std::integral_constant<int, 3>{});
                wait_st();
                fence_before_sync();
                mbar_arrive(&s.bar[DH_FULL + half]);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tbase, bw::TMEM_COLS);
}

int launch_step_bwd_tc(int64_t N, const StepConsts& c, const sphnca_step_params* p, const float* A_s, const float* O_s,
                       const uint8_t* flags_s, const uint32_t* relu_s, const float* dA_next_s, float* dO_s, float* dY_s,
                       float* dAdir_s, cudaStream_t stream, const int32_t* rows, const int32_t* count) {
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        SPHNCA_CUDA(cudaGetDevice(&dev));
        SPHNCA_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int num_tiles = (int)((N + bw::TILE - 1) / bw::TILE);
    const int blocks = num_tiles < num_sms ? num_tiles : num_sms;
    const size_t smem = sizeof(bw::Smem);
    SPHNCA_CUDA(cudaFuncSetAttribute(step_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SPHNCA_LAUNCH(step_bwd_tc_kernel, blocks, bw::THREADS, smem, stream, (int)N, A_s, O_s, flags_s, relu_s, dA_next_s, *p, c,
                  dO_s, dY_s, dAdir_s, num_tiles, rows, count);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

}  // namespace sphnca
