// Weight gradients of the 48->256->n_out MLP on the Blackwell tensor path, from (y, dO) only:
// the hidden layer is RECOMPUTED in the transposed orientation so that the contraction over
// particles lands on the MMA K dimension with the A operand in TMEM:
//
//   recompute   D1t[128h x 64p] = W1_half[128h x 48] . y[64p x 48]^T          (pre-activations^T)
//               D3t[128h x 64p] = W2t_half[128h x 32] . dO[64p x 32]^T        (dL/dH^T before the mask)
//   epilogue    thread = hidden unit: H^T = relu(D1t + b1), dH^T = (D1t + b1 > 0) * (D3t + dO[:,32] W2[32][h]);
//               tf32 hi/lo, written back IN PLACE as TMEM A operands [lanes = h][columns = p]
//   accumulate  dW1_half[128h x 64] += dH^T . [y | 1 | 0]   (column 48 of the result = db1)
//               dW2t_half[128h x 48] += H^T . dO            (B operands = K-major TRANSPOSED copies yT, dOT)
//               The hi and lo B tiles are stacked along N ([y_hi ; y_lo] = 128 rows, [dO_hi ; dO_lo] = 96 rows), so the
//               two passes that share the A operand are ONE MMA: dH_hi . [y_hi ; y_lo] lands in columns 0..63 | 64..127
//               of the accumulator and dH_lo . y_hi adds to columns 64..127 -- 4 MMAs per k-step instead of 6 (an M128 x K8 tf32
//               MMA occupies the tensor pipe for N/2 cycles, probes/tc_timing5.cu: stacking saves instructions and A-operand
//               reads, not pipe time).  The RECOMPUTE MMAs take both operands from shared memory: the 4 KB weight tile (A) alone
//               is 32 cycles of the 128 B/clk port per MMA, 48 with the particle tile, against 33 cycles of pipe time at N = 64
//               -- that GEMM is bound by operand fetch.
//   the two accumulators (224 TMEM columns) live in TMEM; every FLUSH chunks (and at the end) the epilogue warps drain them
//   into the CTA's partial in global memory (+=, fp32 round-to-nearest, owner thread only) and the next accumulate MMA
//   starts from zero: the tensor core's fp32 accumulation TRUNCATES, which biases a long same-sign sum towards zero by
//   ~2e-8 x (number of accumulate steps) -- 3.2e-5 over the 886 steps of a 262,144-row launch, 9e-6 over 256
//   (tests/test_gpu_parity_config_sizes.py::test_weight_gradient_accumulation_262144_rows_vs_fp64).  The drain adds the
//   column groups; one partial per CTA pair (dW1 part stored transposed, [k][h], so that the drains are coalesced) is
//   reduced in fixed order by mlp_wgrad_reduce_kernel (wgrad.cu).
//
// A CTA owns ONE hidden half (blockIdx.x & 1) and a contiguous range of 64-particle chunks, so the
// weights it stages are halves: W1 48 KB + W2^T 32 KB (hi/lo); chunk operand tiles 96 KB.
// No float atomics, fixed order everywhere: deterministic.
#include "step_common.cuh"
#include "tc_common.cuh"

namespace sphnca {

using namespace tc;


namespace wg {
constexpr int CH = 64;                                       // particles per chunk (MMA N / K)
// epilogue warps 0..7: warp & 3 = TMEM lane quarter (a warp may only touch lanes 32*(warp%4)..+31),
// warp >> 2 = which half of the chunk's 64 particle columns it converts
constexpr int EPI_WARPS = 8, MMA_WARP = 8, LOAD_WARP0 = 9, LOAD_WARPS = 4;
constexpr int THREADS = (LOAD_WARP0 + LOAD_WARPS) * 32;      // 416
constexpr int LOADERS = LOAD_WARPS * 32;                     // 128
constexpr int KO = 32;                                       // dO columns on the tensor core in the recompute
constexpr int NY = 64;                                       // rows of yT: 48 inputs + ones + pad
constexpr int NOUT = 48;                                     // rows of dOT: 36 outputs + pad
constexpr int ROW = NF + 2 * NF + NO_PAD;                    // 84 floats per particle: A 16 | Y 32 | dO 36
constexpr int FLUSH = 16;                                    // chunks (1024 rows) accumulated in TMEM between two drains
constexpr int KP = CH + 1;                                   // K-chunk pitch of the K-major chunk tiles in 16-byte units: one
                                                             // pad piece makes the "lanes over k-chunks" writes conflict-free
constexpr uint32_t COL_R1 = 0, COL_HLO = 64, COL_R3 = 128, COL_DHLO = 192, COL_ACC1 = 256, COL_ACC2 = 384, TMEM_COLS = 512;
static_assert(COL_ACC1 + 2 * NY <= COL_ACC2 && COL_ACC2 + 2 * NOUT <= TMEM_COLS, "accumulator columns");
enum { K_FULL = 0, K_EMPTY, T_FULL, T_EMPTY, R_FULL, HT_FULL, HT_FULL_B, HT_FULL_C, HT_FULL_D, ACC_FULL, STG_FULL0, STG_FULL1, NUM_BARS };

struct Smem {
    float W1_hi[128 * NIN], W1_lo[128 * NIN];      // rows = h (this half), K = k
    float W2t_hi[128 * KO], W2t_lo[128 * KO];      // rows = h, K = o
    float yK_hi[KP * NIN], yK_lo[KP * NIN];        // rows = p, K = k        (B of D1t), K-chunk pitch KP*16 B
    float dOK_hi[KP * KO], dOK_lo[KP * KO];        // rows = p, K = o        (B of D3t), K-chunk pitch KP*16 B
    float yT[2 * NY * CH];                         // rows 0..63 = k (+ones) hi, 64..127 = lo ; K = p (B of dW1)
    float dOT[2 * NOUT * CH];                      // rows 0..47 = o hi, 48..95 = lo ;          K = p (B of dW2t)
    // raw fp32 rows of a chunk: three contiguous blocks of the cell-sorted arrays, each filled by ONE TMA bulk
    // copy (double buffered; the 1344 16-byte LDGSTS this replaces took the loader warps ~3100 cycles per chunk)
    float stA[2][CH][NF];
    float stY[2][CH][2 * NF];
    float stD[2][CH][NO_PAD];
    float dO32[2][CH];                             // scalar-gate column dO[:,32] for the epilogue (double buffered)
    float b1[128], w2_last[128];
    uint64_t bar[NUM_BARS];
    uint32_t tmem_base;
};
__device__ __forceinline__ uint32_t fidx(uint32_t rows, uint32_t m, uint32_t k) { return chunk_off(rows, m, k >> 2) / 4 + (k & 3); }
}  // namespace wg

__global__ void __launch_bounds__(wg::THREADS, 1)
mlp_wgrad_tc_kernel(int64_t n_total, const float* __restrict__ A_s, const float* __restrict__ Y_s,
                    const float* __restrict__ dO_s, sphnca_step_params prm, int n_out, float* __restrict__ partial,
                    int part_floats, const int32_t* __restrict__ count_ptr, int acc_in) {
    using namespace wg;
    if (count_ptr != nullptr) n_total = __ldg(count_ptr);     // fire-list mode: the first *count rows of the compact arrays
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int hh = blockIdx.x & 1;                            // hidden half of this CTA
    const int stream_id = blockIdx.x >> 1, num_streams = gridDim.x >> 1;
    const int64_t chunks_total = (n_total + CH - 1) / CH;
    const int64_t c0 = chunks_total * stream_id / num_streams, c1 = chunks_total * (stream_id + 1) / num_streams;
    const int n_mma = n_out < KO ? n_out : KO;

    // W1 [h][k] of this CTA's hidden half -> W1_hi / W1_lo: rows = h, K = k
    stage_chunks<4>(128 * (NIN / 4), tid, THREADS, [&](int idx) {
        int row, ck; stage_map(idx, 128 / 8, row, ck);
        return __ldg(reinterpret_cast<const float4*>(prm.W1 + (size_t)(hh * 128 + row) * NIN) + ck);
    }, [&](int idx, float4 hi, float4 lo) {
        int row, ck; stage_map(idx, 128 / 8, row, ck);
        const uint32_t o = chunk_off(128, row, ck) / 4;
        *reinterpret_cast<float4*>(&s.W1_hi[o]) = hi; *reinterpret_cast<float4*>(&s.W1_lo[o]) = lo;
    });
    // W2 [o][h] -> W2t_hi / W2t_lo: rows = h, K = o
    stage_chunks<3>(128 * (KO / 4), tid, THREADS, [&](int idx) {
        int row, ck; stage_map(idx, 128 / 8, row, ck);
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = 4 * ck + e < n_mma ? __ldg(prm.W2 + (size_t)(4 * ck + e) * NH + hh * 128 + row) : 0.f;
        return make_float4(v[0], v[1], v[2], v[3]);
    }, [&](int idx, float4 hi, float4 lo) {
        int row, ck; stage_map(idx, 128 / 8, row, ck);
        const uint32_t o = chunk_off(128, row, ck) / 4;
        *reinterpret_cast<float4*>(&s.W2t_hi[o]) = hi; *reinterpret_cast<float4*>(&s.W2t_lo[o]) = lo;
    });
    for (int i = tid; i < 128; i += THREADS) {
        s.b1[i] = __ldg(prm.b1 + hh * 128 + i);
        s.w2_last[i] = n_out > KO ? __ldg(prm.W2 + (size_t)KO * NH + hh * 128 + i) : 0.f;
    }
    // constant rows of the transposed B tiles: yT rows 49..63 and dOT rows 36..47 are zero (row 48 of yT,
    // the ones column that yields db1, is rewritten per chunk because of the ragged last chunk)
    for (int i = tid; i < (NY - NIN) * CH; i += THREADS) {
        const int row = NIN + i / CH, pl = i - (i / CH) * CH;
        s.yT[fidx(2 * NY, row, pl)] = 0.f; s.yT[fidx(2 * NY, NY + row, pl)] = 0.f;
    }
    for (int i = tid; i < (NOUT - NO_PAD) * CH; i += THREADS) {
        const int row = NO_PAD + i / CH, pl = i - (i / CH) * CH;
        s.dOT[fidx(2 * NOUT, row, pl)] = 0.f; s.dOT[fidx(2 * NOUT, NOUT + row, pl)] = 0.f;
    }
    if (tid == 0) {
        mbar_init(&s.bar[K_FULL], LOADERS); mbar_init(&s.bar[K_EMPTY], 1);
        mbar_init(&s.bar[T_FULL], LOADERS); mbar_init(&s.bar[T_EMPTY], 1);
        mbar_init(&s.bar[R_FULL], 1);       for (int q = 0; q < 4; ++q) mbar_init(&s.bar[HT_FULL + q], EPI_WARPS * 16);
        mbar_init(&s.bar[ACC_FULL], 1);
        mbar_init(&s.bar[STG_FULL0], 1);    mbar_init(&s.bar[STG_FULL1], 1);
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(&s.tmem_base, TMEM_COLS);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = s.tmem_base;

    if (warp >= LOAD_WARP0) {
        // ============================ loader: chunk operand tiles ============================
        const int lt = tid - LOAD_WARP0 * 32;                 // 0..127
        float db2[NO_PAD];
#pragma unroll
        for (int o = 0; o < NO_PAD; ++o) db2[o] = 0.f;
        // staging: one elected loader thread issues the three bulk copies of a chunk a full chunk ahead; rows past
        // the end of the arrays (ragged last chunk) are zero-filled by plain stores
        auto issue_copy = [&](int64_t chn, int buf) {
            const int64_t p0 = chn * CH;
            const int nv = (int)((n_total - p0) < CH ? (n_total - p0) : CH);
            if (nv < CH) {
                for (int i = lt; i < (CH - nv) * ROW; i += LOADERS) {
                    const int pl = nv + i / ROW, c = i - (i / ROW) * ROW;
                    if (c < NF) s.stA[buf][pl][c] = 0.f;
                    else if (c < 3 * NF) s.stY[buf][pl][c - NF] = 0.f;
                    else s.stD[buf][pl][c - 3 * NF] = 0.f;
                }
            }
            if (lt == 0) {
                uint64_t* bar = &s.bar[STG_FULL0 + buf];
                mbar_arrive_expect_tx(bar, (uint32_t)nv * ROW * 4u);
                bulk_g2s(&s.stA[buf][0][0], A_s + p0 * NF, (uint32_t)nv * NF * 4u, bar);
                bulk_g2s(&s.stY[buf][0][0], Y_s + p0 * 2 * NF, (uint32_t)nv * 2 * NF * 4u, bar);
                bulk_g2s(&s.stD[buf][0][0], dO_s + p0 * NO_PAD, (uint32_t)nv * NO_PAD * 4u, bar);
            }
        };
        auto split4 = [](const float4 v, float4& hi, float4& lo) {
            split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
        };
        if (c0 < c1) issue_copy(c0, 0);
        int it = 0;
        for (int64_t ch = c0; ch < c1; ++ch, ++it) {
            const int buf = it & 1;
            // the other staging buffer was last read in phase 2 of the previous chunk (all loaders passed the
            // barrier at the end of that iteration): refill it now, a full chunk ahead
            if (lt == 0) SPHNCA_TL(13);
            if (ch + 1 < c1) issue_copy(ch + 1, buf ^ 1);
            if (lt == 0) SPHNCA_TL(14);
            mbar_wait(&s.bar[STG_FULL0 + buf], (it >> 1) & 1);           // this chunk's rows have landed
            if (lt == 0) SPHNCA_TL(11);
            if ((ch + 1) * CH > n_total)                                  // ragged last chunk: the zero-filled tail rows must be visible
                asm volatile("bar.sync 1, %0;" :: "r"(LOADERS) : "memory");
            if (lt == 0) SPHNCA_TL(12);
            // phase 1 (K-major tiles): a piece = 4 consecutive k of one particle.  Lanes run over the pieces of a
            // row (contiguous 16 B reads); the tile's K-chunk pitch is KP*16 B, so those lanes hit distinct banks.
            mbar_wait(&s.bar[K_EMPTY], (it & 1) ^ 1);         // recompute MMAs of the previous chunk are done
            if (lt == 0) SPHNCA_TL(0);
#pragma unroll
            for (int r = 0; r < CH * (NF / 4) / LOADERS; ++r) {           // A part: k chunks 0..3
                const int i = lt + r * LOADERS, pl = i >> 2, q = i & 3;
                float4 hi, lo; split4(*reinterpret_cast<const float4*>(&s.stA[buf][pl][4 * q]), hi, lo);
                const uint32_t o = (q * KP + pl) * 4;
                *reinterpret_cast<float4*>(&s.yK_hi[o]) = hi; *reinterpret_cast<float4*>(&s.yK_lo[o]) = lo;
            }
#pragma unroll
            for (int r = 0; r < CH * (2 * NF / 4) / LOADERS; ++r) {       // Y part: k chunks 4..11
                const int i = lt + r * LOADERS, pl = i >> 3, q = i & 7;
                float4 hi, lo; split4(*reinterpret_cast<const float4*>(&s.stY[buf][pl][4 * q]), hi, lo);
                const uint32_t o = ((q + NF / 4) * KP + pl) * 4;
                *reinterpret_cast<float4*>(&s.yK_hi[o]) = hi; *reinterpret_cast<float4*>(&s.yK_lo[o]) = lo;
            }
#pragma unroll
            for (int r = 0; r < CH * (KO / 4) / LOADERS; ++r) {           // dO columns 0..31
                const int i = lt + r * LOADERS, pl = i >> 3, q = i & 7;
                float4 hi, lo; split4(*reinterpret_cast<const float4*>(&s.stD[buf][pl][4 * q]), hi, lo);
                const uint32_t o = (q * KP + pl) * 4;
                *reinterpret_cast<float4*>(&s.dOK_hi[o]) = hi; *reinterpret_cast<float4*>(&s.dOK_lo[o]) = lo;
            }
            if (lt < CH) s.dO32[it & 1][lt] = s.stD[buf][lt][KO];
            fence_proxy_async_smem();
            mbar_arrive(&s.bar[K_FULL]);
            if (lt == 0) SPHNCA_TL(1);
            if (hh == 0 && lt < CH) {                         // db2 partial sums (one CTA of the pair)
#pragma unroll
                for (int o = 0; o < NO_PAD; ++o) db2[o] += s.stD[buf][lt][o];
            }
            // phase 2 (transposed K-major tiles): one 16-byte piece = 4 consecutive particles of one row.
            // Lanes vary over rows: consecutive words of a staged row, consecutive 16 B of the tile.
            mbar_wait(&s.bar[T_EMPTY], (it & 1) ^ 1);         // accumulate MMAs of the previous chunk are done
            if (lt == 0) SPHNCA_TL(2);
            // three convergent loops (one per staged block); in the A block a warp covers two particle groups whose
            // rows are 64 words apart, so its upper half reads the 4 particles in the order 1,0,3,2 (distinct banks)
            auto put = [&](float* t, uint32_t o, uint32_t lo_rows, const float4 v) {     // hi piece at o, lo piece lo_rows below
                float4 hi, lo; split4(v, hi, lo);
                *reinterpret_cast<float4*>(&t[o]) = hi; *reinterpret_cast<float4*>(&t[o + 4 * lo_rows]) = lo;
            };
#pragma unroll
            for (int r = 0; r < NF * (CH / 4) / LOADERS; ++r) {           // A rows 0..15
                const int i = lt + r * LOADERS, pg = i >> 4, row = i & 15, sw = (lane >> 4) & 1;
                const float a = s.stA[buf][4 * pg + sw][row], b = s.stA[buf][4 * pg + (sw ^ 1)][row];
                const float c = s.stA[buf][4 * pg + 2 + sw][row], d = s.stA[buf][4 * pg + 2 + (sw ^ 1)][row];
                put(s.yT, chunk_off(2 * NY, row, pg) / 4, NY, sw ? make_float4(b, a, d, c) : make_float4(a, b, c, d));
            }
#pragma unroll
            for (int r = 0; r < 2 * NF * (CH / 4) / LOADERS; ++r) {       // Y rows 16..47
                const int i = lt + r * LOADERS, pg = i >> 5, row = i & 31;
                put(s.yT, chunk_off(2 * NY, NF + row, pg) / 4, NY,
                    make_float4(s.stY[buf][4 * pg][row], s.stY[buf][4 * pg + 1][row], s.stY[buf][4 * pg + 2][row], s.stY[buf][4 * pg + 3][row]));
            }
#pragma unroll
            for (int r = 0; r < (NO_PAD * (CH / 4) + LOADERS - 1) / LOADERS; ++r) {   // dO rows 0..35
                const int i = lt + r * LOADERS;
                if (i < NO_PAD * (CH / 4)) {
                    const int pg = i / NO_PAD, row = i - pg * NO_PAD;
                    put(s.dOT, chunk_off(2 * NOUT, row, pg) / 4, NOUT,
                        make_float4(s.stD[buf][4 * pg][row], s.stD[buf][4 * pg + 1][row], s.stD[buf][4 * pg + 2][row], s.stD[buf][4 * pg + 3][row]));
                }
            }
            if (lt < CH / 4) {                                // ones row (-> db1): 1 for real particles
                float one[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) one[j] = (ch * CH + 4 * lt + j < n_total) ? 1.f : 0.f;
                *reinterpret_cast<float4*>(&s.yT[chunk_off(2 * NY, NIN, lt) / 4]) = make_float4(one[0], one[1], one[2], one[3]);
            }
            fence_proxy_async_smem();
            mbar_arrive(&s.bar[T_FULL]);
            if (lt == 0) SPHNCA_TL(3);
            // every loader is done reading this staging buffer before it is refilled two iterations later
            asm volatile("bar.sync 1, %0;" :: "r"(LOADERS) : "memory");
        }
        // db2: fixed-order tree over the 64 row-owning loader threads, written by the hh == 0 CTA
        if (hh == 0) {
            // scratch for the tree: the dO staging buffer (idle now)
            float (*red)[NO_PAD] = s.stD[0];
            if (lt < CH) {
#pragma unroll
                for (int o = 0; o < NO_PAD; ++o) red[lt][o] = db2[o];
            }
            asm volatile("bar.sync 1, %0;" :: "r"(LOADERS) : "memory");
            for (int stride = CH / 2; stride > 0; stride >>= 1) {
                if (lt < stride) {
#pragma unroll
                    for (int o = 0; o < NO_PAD; ++o) red[lt][o] += red[lt + stride][o];
                }
                asm volatile("bar.sync 1, %0;" :: "r"(LOADERS) : "memory");
            }
            if (lt < NO_PAD) {                               // acc_in: the partials of earlier launches stay (one reduction per rollout)
                float* d = partial + (size_t)stream_id * part_floats + NH * NIN + NH + NO_PAD * NH + lt;
                *d = acc_in ? *d + red[0][lt] : red[0][lt];
            }
        }
    } else if (warp == MMA_WARP) {
        {   // warp-uniform issue: every lane runs the loop, one elected lane issues (tc_common.cuh)
            const uint32_t go = elect_one();
            const uint32_t idR = make_idesc_tf32(128, CH), idA1 = make_idesc_tf32(128, NY), idA2 = make_idesc_tf32(128, NOUT);
            const uint32_t idA1c = make_idesc_tf32(128, 2 * NY), idA2c = make_idesc_tf32(128, 2 * NOUT);   // stacked [hi ; lo] B tiles
            // base descriptors, built once (desc_advance, tc_common.cuh)
            const uint64_t dW1h = make_smem_desc(smem_u32(s.W1_hi), 128 * 16, 128), dW1l = make_smem_desc(smem_u32(s.W1_lo), 128 * 16, 128);
            const uint64_t dW2h = make_smem_desc(smem_u32(s.W2t_hi), 128 * 16, 128), dW2l = make_smem_desc(smem_u32(s.W2t_lo), 128 * 16, 128);
            const uint64_t dyKh = make_smem_desc(smem_u32(s.yK_hi), KP * 16, 128), dyKl = make_smem_desc(smem_u32(s.yK_lo), KP * 16, 128);
            const uint64_t dOKh = make_smem_desc(smem_u32(s.dOK_hi), KP * 16, 128), dOKl = make_smem_desc(smem_u32(s.dOK_lo), KP * 16, 128);
            const uint64_t dyT = make_smem_desc(smem_u32(s.yT), 2 * NY * 16, 128), dOT = make_smem_desc(smem_u32(s.dOT), 2 * NOUT * 16, 128);
            int it = 0;
            for (int64_t ch = c0; ch < c1; ++ch, ++it) {
                mbar_wait(&s.bar[K_FULL], it & 1);
                fence_after_sync();
                SPHNCA_TL(4);
                // recompute (the TMEM chunk regions were last read by the accumulate MMAs issued before: in order)
#pragma unroll
                for (int ks = 0; ks < NIN / 8; ++ks) {
                    const uint64_t a_hi = desc_advance(dW1h, chunk_off(128, 0, 2 * ks)), a_lo = desc_advance(dW1l, chunk_off(128, 0, 2 * ks));
                    const uint64_t b_hi = desc_advance(dyKh, 2 * ks * KP * 16), b_lo = desc_advance(dyKl, 2 * ks * KP * 16);
                    mma_tf32_ss_e(tbase + COL_R1, a_lo, b_hi, idR, ks > 0, go);
                    mma_tf32_ss_e(tbase + COL_R1, a_hi, b_lo, idR, true, go);
                    mma_tf32_ss_e(tbase + COL_R1, a_hi, b_hi, idR, true, go);
                }
#pragma unroll
                for (int ks = 0; ks < KO / 8; ++ks) {
                    const uint64_t a_hi = desc_advance(dW2h, chunk_off(128, 0, 2 * ks)), a_lo = desc_advance(dW2l, chunk_off(128, 0, 2 * ks));
                    const uint64_t b_hi = desc_advance(dOKh, 2 * ks * KP * 16), b_lo = desc_advance(dOKl, 2 * ks * KP * 16);
                    mma_tf32_ss_e(tbase + COL_R3, a_lo, b_hi, idR, ks > 0, go);
                    mma_tf32_ss_e(tbase + COL_R3, a_hi, b_lo, idR, true, go);
                    mma_tf32_ss_e(tbase + COL_R3, a_hi, b_hi, idR, true, go);
                }
                mma_commit_e(&s.bar[R_FULL], go);
                mma_commit_e(&s.bar[K_EMPTY], go);
                SPHNCA_TL(5);
                mbar_wait(&s.bar[HT_FULL], it & 1);
                SPHNCA_TL(6);
                mbar_wait(&s.bar[T_FULL], it & 1);
                fence_after_sync();
                SPHNCA_TL(7);
                const bool fresh = (it % FLUSH) == 0;        // the accumulators were drained (or never written): start from zero
                if (fresh && it > 0) {                        // both warp groups drained before they converted their first quarter
                    mbar_wait(&s.bar[HT_FULL + 1], it & 1);
                    fence_after_sync();
                }
#pragma unroll
                for (int ks = 0; ks < CH / 8; ++ks) {
                    if (ks > 0 && (ks & 1) == 0) { mbar_wait(&s.bar[HT_FULL + ks / 2], it & 1); fence_after_sync(); }   // next 16 columns converted
                    // one descriptor per tile: rows 0..N-1 of it are the hi tile, all 2N rows the stacked [hi ; lo] tile
                    const uint64_t y_t = desc_advance(dyT, chunk_off(2 * NY, 0, 2 * ks)), o_t = desc_advance(dOT, chunk_off(2 * NOUT, 0, 2 * ks));
                    const bool acc = !(fresh && ks == 0);
                    mma_tf32_ts_e(tbase + COL_ACC1, tbase + COL_R3 + 8 * ks, y_t, idA1c, acc, go);      // dH_hi . [y_hi ; y_lo]
                    mma_tf32_ts_e(tbase + COL_ACC1 + NY, tbase + COL_DHLO + 8 * ks, y_t, idA1, true, go);   // dH_lo . y_hi -> second group: the first one
                                                                                                           // holds the large hi . hi terms only (half the truncating adds)
                    mma_tf32_ts_e(tbase + COL_ACC2, tbase + COL_R1 + 8 * ks, o_t, idA2c, acc, go);      // H_hi . [dO_hi ; dO_lo]
                    mma_tf32_ts_e(tbase + COL_ACC2 + NOUT, tbase + COL_HLO + 8 * ks, o_t, idA2, true, go);  // H_lo . dO_hi -> second group
                }
                mma_commit_e(&s.bar[T_EMPTY], go);
                SPHNCA_TL(8);
            }
            mma_commit_e(&s.bar[ACC_FULL], go);
        }
    } else {
        // ============================ epilogue: thread = hidden unit ============================
        const int h = (warp & 3) * 32 + lane;                 // 0..127 within the half == TMEM lane
        const int col0 = (warp >> 2) * (CH / 4);              // first of this warp's two 16-column quarters
        const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        const float b1h = s.b1[h], w2l = s.w2_last[h];
        // drain of the accumulators into this CTA pair's partial: warp group 0 -> dW1 row h (48) | db1 (column 48),
        // warp group 1 -> dW2[o][h] (column o of ACC2).  first = nothing drained yet (plain store), else +=.
        float* out = partial + (size_t)stream_id * part_floats;
        const int hg = hh * 128 + h;
        auto drain = [&](const bool first) {
            if (warp < 4) {
#pragma unroll
                for (int cc = 0; cc < NY; cc += 16) {
                    uint32_t r[16], r2[16];                  // columns of the hi.(hi + lo of A) group and of the hi.lo group
                    tmem_ld16(lane_addr + COL_ACC1 + cc, r); tmem_ld16(lane_addr + COL_ACC1 + NY + cc, r2);
                    float prev[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int k = cc + j;
                        prev[j] = (first || k > NIN) ? 0.f : out[k * NH + hg];        // dW1 TRANSPOSED [k][h] | db1 right behind it
                    }
                    wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int k = cc + j;
                        const float val = prev[j] + (__uint_as_float(r[j]) + __uint_as_float(r2[j]));
                        if (k <= NIN) out[k * NH + hg] = val;       // lanes = consecutive h: coalesced (148 CTAs drain in lockstep)
                    }
                }
            } else {
#pragma unroll
                for (int cc = 0; cc < NOUT; cc += 16) {
                    uint32_t r[16], r2[16];
                    tmem_ld16(lane_addr + COL_ACC2 + cc, r); tmem_ld16(lane_addr + COL_ACC2 + NOUT + cc, r2);
                    float prev[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int o = cc + j;
                        prev[j] = (first || o >= NO_PAD) ? 0.f : out[NH * NIN + NH + o * NH + hg];
                    }
                    wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int o = cc + j;
                        if (o < NO_PAD) out[NH * NIN + NH + o * NH + hg] = prev[j] + (__uint_as_float(r[j]) + __uint_as_float(r2[j]));
                    }
                }
            }
        };
        int it = 0;
        for (int64_t ch = c0; ch < c1; ++ch, ++it) {
            mbar_wait(&s.bar[R_FULL], it & 1);
            fence_after_sync();
            if (tid == 0) SPHNCA_TL(9);
            // R_FULL(it) follows the accumulate MMAs of chunk it-1 in issue order: the accumulators are final up to there
            if (it > 0 && (it % FLUSH) == 0) drain(it == FLUSH && !acc_in);
            // 16 particle columns per quarter; warp group g = warp >> 2 converts quarters g and g + 2, so quarters 0|1 and
            // 2|3 are produced in parallel and the accumulate MMAs over the first particles start while the later columns
            // are still being converted; the second quarter's TMEM loads are in flight during the first one's arithmetic
            uint32_t r1[2][16], r3[2][16];
            tmem_ld16(lane_addr + COL_R1 + col0, r1[0]);
            tmem_ld16(lane_addr + COL_R3 + col0, r3[0]);
#pragma unroll
            for (int ph = 0; ph < 2; ++ph) {
                const int qt = 2 * ph + (warp >> 2);
                const int cc = qt * (CH / 4);
                uint32_t hhi[16], hlo[16], dhi[16], dlo[16];
                wait_ld();
                if (ph == 0) {
                    tmem_ld16(lane_addr + COL_R1 + cc + CH / 2, r1[1]);
                    tmem_ld16(lane_addr + COL_R3 + cc + CH / 2, r3[1]);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float pre = __uint_as_float(r1[ph][j]) + b1h;
                    const bool on = pre > 0.f;
                    const float hv = on ? pre : 0.f;
                    const float dh = on ? fmaf(s.dO32[it & 1][cc + j], w2l, __uint_as_float(r3[ph][j])) : 0.f;
                    float a, b;
                    split_tf32(hv, a, b); hhi[j] = __float_as_uint(a); hlo[j] = __float_as_uint(b);
                    split_tf32(dh, a, b); dhi[j] = __float_as_uint(a); dlo[j] = __float_as_uint(b);
                }
                tmem_st16(lane_addr + COL_R1 + cc, hhi);
                tmem_st16(lane_addr + COL_HLO + cc, hlo);
                tmem_st16(lane_addr + COL_R3 + cc, dhi);
                tmem_st16(lane_addr + COL_DHLO + cc, dlo);
                wait_st();
                fence_before_sync();
                mbar_arrive(&s.bar[HT_FULL + qt]);
            }
            if (tid == 0) SPHNCA_TL(10);
        }
        if (c0 < c1) {
            mbar_wait(&s.bar[ACC_FULL], 0);
            fence_after_sync();
            drain(it <= FLUSH && !acc_in);                   // the chunks since the last drain (at least one)
        } else if (warp < 4 && !acc_in) {                    // a stream without chunks: zeros
            for (int k = 0; k <= NIN; ++k) out[k * NH + hg] = 0.f;
            for (int o = 0; o < NO_PAD; ++o) out[NH * NIN + NH + o * NH + hg] = 0.f;
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tbase, wg::TMEM_COLS);
}

int wgrad_tc_streams(int64_t N) {
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) num_sms = 148;
    }
    int64_t chunks = (N + wg::CH - 1) / wg::CH;
    int64_t st = num_sms / 2;
    if (st > chunks) st = chunks;
    return (int)(st < 1 ? 1 : st);
}

int launch_mlp_wgrad_tc(int64_t N, const float* A_s, const float* Y_s, const float* dO_s, const sphnca_step_params* p,
                        int n_out, float* partial, int part_floats, int streams, cudaStream_t stream, const int32_t* count, int acc_in) {
    const size_t smem = sizeof(wg::Smem);
    SPHNCA_CUDA(cudaFuncSetAttribute(mlp_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SPHNCA_LAUNCH(mlp_wgrad_tc_kernel, 2 * streams, wg::THREADS, smem, stream, N, A_s, Y_s, dO_s, *p, n_out, partial, part_floats, count, acc_in);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

}  // namespace sphnca

SPHNCA_TIMELINE_READER(sphnca_debug_timeline_wgrad)
