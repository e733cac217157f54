// Fused NCA step, forward, on the Blackwell tensor path (sm_100a): persistent, warp-specialised.
//
//   sweep warps (16)   4 lanes per particle walk the 3^D-cell neighbourhood of a 128-particle tile,
//                      reduce SPH gradient (x,y; 16 channels) + alive-bit blur with warp shuffles and
//                      write the 48-wide perception row STRAIGHT into the MMA A tile in shared memory
//                      (tf32 hi/lo split, K-major core-matrix layout) -- it never round-trips HBM.
//   MMA warp (1 lane)  tcgen05.mma kind::tf32, 3xTF32 split for fp32 parity:
//                        GEMM1  D1[128 x 128] = Y[128x48] . W1half^T      (twice: hidden 0..127, 128..255)
//                        GEMM2  D2[128 x 32] += H[128x128] . W2half^T     (A operand = H in TMEM)
//                      accumulators in TMEM, completion signalled with tcgen05.commit -> mbarrier.
//   epilogue warps (4) tcgen05.ld D1 -> +b1, ReLU -> tf32 hi/lo -> tcgen05.st H (feeds GEMM2 from TMEM);
//                      the 33rd MLP output (the scalar gate) is a 256-long dot product done here on
//                      the fly, so GEMM2 is N=32; final: tcgen05.ld D2 -> gated/orig update + fire mask.
//
// Shared memory (215 KB of 227 KB): W1 hi/lo 96 KB + W2[0..31] hi/lo 64 KB staged once per CTA,
// one Y tile hi/lo 48 KB.  TMEM: D1 128 + H_hi 128 + H_lo 128 + D2 32 = 416 of 512 columns.
// One CTA per SM, grid = min(#tiles, #SMs); tiles are taken round-robin.
//
// Produces exactly the outputs of step_fwd_kernel (step.cu): A_new, alive, Y, O, flags.
#include "step_common.cuh"
#include "tc_common.cuh"

namespace sphnca {

using namespace tc;

constexpr int TILE = 128;
constexpr int SWEEP_LANES = 4;
constexpr int EPI_WARPS = 4;
constexpr int MMA_WARP = EPI_WARPS;
constexpr int SWEEP_WARP0 = EPI_WARPS + 1;
constexpr int SWEEP_WARPS = TILE * SWEEP_LANES / 32;
constexpr int TC_THREADS = (SWEEP_WARP0 + SWEEP_WARPS) * 32;   // 672
constexpr int N2 = 32;                                          // GEMM2 width (outputs 0..31)
constexpr uint32_t COL_D1 = 0, COL_HHI = 128, COL_HLO = 256, COL_D2 = 384, TMEM_COLS = 512;

enum { Y_FULL = 0, Y_EMPTY, D1_FULL, D1_EMPTY, H_FULL, H_EMPTY, D2_FULL, D2_EMPTY, NUM_BARS };

struct TcSmem {
    float W1_hi[2][128 * NIN], W1_lo[2][128 * NIN];   // [hidden half][chunk layout of 128 rows x 48]
    float W2_hi[2][N2 * 128], W2_lo[2][N2 * 128];     // [k half][chunk layout of 32 rows x 128]
    float Y_hi[TILE * NIN], Y_lo[TILE * NIN];         // chunk layout of 128 rows x 48
    float b1[NH];
    float w2_last[NH];                                // row 32 of W2 (the scalar gate), fp32
    float b2[NO_PAD];
    uint8_t prev[2][TILE];
    uint64_t bar[NUM_BARS];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t fidx(uint32_t rows, uint32_t m, uint32_t k) {   // float index in a chunk-layout tile
    return chunk_off(rows, m, k >> 2) / 4 + (k & 3);
}

template <int D>
__global__ void __launch_bounds__(TC_THREADS, 1)
step_fwd_tc_kernel(GridView g, const float4* __restrict__ xv4, const float* __restrict__ A_s,
                   const uint8_t* __restrict__ fire_s, sphnca_step_params prm, StepConsts c,
                   float* __restrict__ An_s, float* __restrict__ Y_s, float* __restrict__ O_s,
                   uint8_t* __restrict__ flags_s, uint8_t* __restrict__ alive_s, uint32_t* __restrict__ relu_s,
                   int num_tiles, int dbg) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TcSmem& s = *reinterpret_cast<TcSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup: weights -> tf32 hi/lo operand tiles, barriers, TMEM ----
    for (int i = tid; i < NH * NIN; i += TC_THREADS) {
        const int n = i / NIN, k = i - n * NIN;
        float hi, lo; split_tf32(__ldg(prm.W1 + i), hi, lo);
        const uint32_t o = fidx(128, n & 127, k);
        s.W1_hi[n >> 7][o] = hi; s.W1_lo[n >> 7][o] = lo;
    }
    const int n_mma = c.n_out < N2 ? c.n_out : N2;
    for (int i = tid; i < N2 * NH; i += TC_THREADS) {
        const int o_ = i / NH, k = i - o_ * NH;
        float hi, lo; split_tf32(o_ < n_mma ? __ldg(prm.W2 + (size_t)o_ * NH + k) : 0.f, hi, lo);
        const uint32_t o = fidx(N2, o_, k & 127);
        s.W2_hi[k >> 7][o] = hi; s.W2_lo[k >> 7][o] = lo;
    }
    for (int i = tid; i < NH; i += TC_THREADS) {
        s.b1[i] = __ldg(prm.b1 + i);
        s.w2_last[i] = c.n_out > N2 ? __ldg(prm.W2 + (size_t)N2 * NH + i) : 0.f;
    }
    if (tid < NO_PAD) s.b2[tid] = tid < c.n_out ? __ldg(prm.b2 + tid) : 0.f;
    if (tid == 0) {
        mbar_init(&s.bar[Y_FULL], SWEEP_WARPS * 32); mbar_init(&s.bar[Y_EMPTY], 1);
        mbar_init(&s.bar[D1_FULL], 1);               mbar_init(&s.bar[D1_EMPTY], EPI_WARPS * 32);
        mbar_init(&s.bar[H_FULL], EPI_WARPS * 32);   mbar_init(&s.bar[H_EMPTY], 1);
        mbar_init(&s.bar[D2_FULL], 1);               mbar_init(&s.bar[D2_EMPTY], EPI_WARPS * 32);
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(&s.tmem_base, TMEM_COLS);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = s.tmem_base;

    if (warp >= SWEEP_WARP0) {
        // =========================== perception sweep (producer) ===========================
        const int st = tid - SWEEP_WARP0 * 32;
        const int m = st >> 2, sub = st & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int p = tile * TILE + m;
            const bool valid = p < g.n;
            if (st == 0) SPHNCA_TL(0);
            float Ai[NF], gx[NF], gy[NF];
            float blur = 0.f;
#pragma unroll
            for (int f = 0; f < NF; ++f) { Ai[f] = 0.f; gx[f] = 0.f; gy[f] = 0.f; }
            if (valid && dbg != 2) {
                const float4 xi = xv4[p];
                load16(A_s + (size_t)p * NF, Ai);
                for_each_run<D>(g, g.sorted_cell[p], [&](int q0, int q1) {
                    for (int q = q0 + sub; q < q1; q += SWEEP_LANES) {
                        float rx, ry;
                        const float4 xj = __ldg(&xv4[q]);
                        const float d2 = pair_d2<D>(xi, xj, rx, ry);
                        if (d2 < c.h2) {
                            float Aj[NF];
                            load16(A_s + (size_t)q * NF, Aj);
                            const float t = c.h2 - d2;
                            const float alive = c.use_alpha ? (Aj[3] > 0.1f ? 1.f : 0.f) : 1.f;
                            blur = fmaf(alive * xj.w, t * t * t, blur);
                            if (d2 > 0.f) {
                                const float d = sqrtf(d2);
                                const float hd = c.h - d;
                                const float mg = 3.f * hd * hd / d * xj.w;
                                const float wx = mg * rx, wy = mg * ry;
#pragma unroll
                                for (int f = 0; f < NF; ++f) {
                                    const float dA = Aj[f] - Ai[f];
                                    gx[f] = fmaf(dA, wx, gx[f]);
                                    gy[f] = fmaf(dA, wy, gy[f]);
                                }
                            }
                        }
                    }
                });
            }
            if (st == 0) SPHNCA_TL(1);
            // butterfly over the 4 lanes of a particle: every lane ends with the full sums
#pragma unroll
            for (int o = 1; o < SWEEP_LANES; o <<= 1) {
                blur += __shfl_xor_sync(0xffffffffu, blur, o);
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    gx[f] += __shfl_xor_sync(0xffffffffu, gx[f], o);
                    gy[f] += __shfl_xor_sync(0xffffffffu, gy[f], o);
                }
            }
#pragma unroll
            for (int f = 0; f < NF; ++f) { gx[f] *= c.kg; gy[f] *= c.kg; }

            mbar_wait(&s.bar[Y_EMPTY], (it & 1) ^ 1);      // GEMM1 of the previous tile has consumed Y
            if (st == 0) SPHNCA_TL(2);
            // perception row y = [A | gx | gy] = 12 chunks of 4; lane `sub` writes chunks 3*sub .. 3*sub+2
#pragma unroll
            for (int ch = 0; ch < NIN / 4; ++ch) {
                if (ch / 3 == sub) {
                    float v[4], hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = 4 * ch + j;
                        v[j] = k < NF ? Ai[k] : (k < 2 * NF ? gx[k - NF] : gy[k - 2 * NF]);
                        split_tf32(v[j], hi[j], lo[j]);
                    }
                    const uint32_t o = chunk_off(TILE, m, ch) / 4;
                    *reinterpret_cast<float4*>(&s.Y_hi[o]) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(&s.Y_lo[o]) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            if (valid) {
                // saved perception (for backward), 8 float4 per particle: lane `sub` writes 2 of them
                float4* y4 = reinterpret_cast<float4*>(Y_s + (size_t)p * 2 * NF);
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    if ((k4 >> 1) == sub) {
                        const int k = 4 * k4;
                        y4[k4] = k < NF ? make_float4(gx[k], gx[k + 1], gx[k + 2], gx[k + 3])
                                        : make_float4(gy[k - NF], gy[k - NF + 1], gy[k - NF + 2], gy[k - NF + 3]);
                    }
                }
            }
            if (sub == 0) s.prev[it & 1][m] = (blur * c.cw > 0.1f) ? 1 : 0;
            fence_proxy_async_smem();
            mbar_arrive(&s.bar[Y_FULL]);
            if (st == 0) SPHNCA_TL(3);
        }
    } else if (warp == MMA_WARP) {
        // =========================== tensor-core issue (one lane) ===========================
        if (lane == 0) {
            const uint32_t idesc1 = make_idesc_tf32(TILE, 128), idesc2 = make_idesc_tf32(TILE, N2);
            const uint32_t yh = smem_u32(s.Y_hi), yl = smem_u32(s.Y_lo);
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                mbar_wait(&s.bar[Y_FULL], it & 1);
                fence_after_sync();
                SPHNCA_TL(4);
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    if (half == 1) { mbar_wait(&s.bar[D1_EMPTY], 0); fence_after_sync(); SPHNCA_TL(6); }
                    else if (it > 0) { mbar_wait(&s.bar[D1_EMPTY], 1); fence_after_sync(); }
                    const uint32_t wh = smem_u32(s.W1_hi[half]), wl = smem_u32(s.W1_lo[half]);
#pragma unroll
                    for (int ks = 0; ks < NIN / 8; ++ks) {
                        const uint64_t a_hi = make_smem_desc(yh + chunk_off(TILE, 0, 2 * ks), TILE * 16, 128);
                        const uint64_t a_lo = make_smem_desc(yl + chunk_off(TILE, 0, 2 * ks), TILE * 16, 128);
                        const uint64_t b_hi = make_smem_desc(wh + chunk_off(128, 0, 2 * ks), 128 * 16, 128);
                        const uint64_t b_lo = make_smem_desc(wl + chunk_off(128, 0, 2 * ks), 128 * 16, 128);
                        mma_tf32_ss(tbase + COL_D1, a_lo, b_hi, idesc1, ks > 0);
                        mma_tf32_ss(tbase + COL_D1, a_hi, b_lo, idesc1, true);
                        mma_tf32_ss(tbase + COL_D1, a_hi, b_hi, idesc1, true);
                    }
                    mma_commit(&s.bar[D1_FULL]);
                    SPHNCA_TL(5 + 2 * half);
                }
                mma_commit(&s.bar[Y_EMPTY]);                  // all GEMM1 reads of the Y tile are done
                mbar_wait(&s.bar[D2_EMPTY], (it & 1) ^ 1);    // epilogue drained D2 of the previous tile
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    mbar_wait(&s.bar[H_FULL], half);
                    fence_after_sync();
                    SPHNCA_TL(8 + 2 * half);
                    const uint32_t wh = smem_u32(s.W2_hi[half]), wl = smem_u32(s.W2_lo[half]);
#pragma unroll
                    for (int ks = 0; ks < 128 / 8; ++ks) {
                        const uint64_t b_hi = make_smem_desc(wh + chunk_off(N2, 0, 2 * ks), N2 * 16, 128);
                        const uint64_t b_lo = make_smem_desc(wl + chunk_off(N2, 0, 2 * ks), N2 * 16, 128);
                        mma_tf32_ts(tbase + COL_D2, tbase + COL_HLO + 8 * ks, b_hi, idesc2, (half | ks) > 0);
                        mma_tf32_ts(tbase + COL_D2, tbase + COL_HHI + 8 * ks, b_lo, idesc2, true);
                        mma_tf32_ts(tbase + COL_D2, tbase + COL_HHI + 8 * ks, b_hi, idesc2, true);
                    }
                    mma_commit(&s.bar[H_EMPTY]);
                    SPHNCA_TL(9 + 2 * half);
                }
                mma_commit(&s.bar[D2_FULL]);
            }
        }
    } else {
        // =========================== epilogue (TMEM -> registers) ===========================
        const int m = tid;                                   // tile row == TMEM lane
        const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int p = tile * TILE + m;
            const bool valid = p < g.n;
            float o32 = 0.f;
            uint32_t rmask[NH / 32];
#pragma unroll
            for (int wd = 0; wd < NH / 32; ++wd) rmask[wd] = 0u;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                mbar_wait(&s.bar[D1_FULL], half);
                mbar_wait(&s.bar[H_EMPTY], half ^ 1);        // GEMM2 finished with the previous H
                fence_after_sync();
                if (tid == 0) SPHNCA_TL(12 + half);
#pragma unroll
                for (int cc = 0; cc < 128; cc += 16) {
                    uint32_t r[16], hi[16], lo[16];
                    tmem_ld16(lane_addr + COL_D1 + cc, r);
                    wait_ld();
                    uint32_t bits = 0;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = half * 128 + cc + j;
                        const float pre = __uint_as_float(r[j]) + s.b1[n];
                        bits |= (pre > 0.f ? 1u : 0u) << j;
                        const float hv = fmaxf(pre, 0.f);
                        o32 = fmaf(hv, s.w2_last[n], o32);
                        float h_, l_; split_tf32(hv, h_, l_);
                        hi[j] = __float_as_uint(h_); lo[j] = __float_as_uint(l_);
                    }
                    rmask[(half * 128 + cc) >> 5] |= bits << (cc & 16);
                    tmem_st16(lane_addr + COL_HHI + cc, hi);
                    tmem_st16(lane_addr + COL_HLO + cc, lo);
                }
                wait_st();
                fence_before_sync();
                mbar_arrive(&s.bar[D1_EMPTY]);
                mbar_arrive(&s.bar[H_FULL]);
            }
            mbar_wait(&s.bar[D2_FULL], it & 1);
            fence_after_sync();
            if (tid == 0) SPHNCA_TL(14);
            float out[NO_PAD];
            {
                uint32_t r[16];
                tmem_ld16(lane_addr + COL_D2, r); wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) out[j] = __uint_as_float(r[j]) + s.b2[j];
                tmem_ld16(lane_addr + COL_D2 + 16, r); wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) out[16 + j] = __uint_as_float(r[j]) + s.b2[16 + j];
            }
            fence_before_sync();
            mbar_arrive(&s.bar[D2_EMPTY]);
            out[32] = o32 + s.b2[32]; out[33] = 0.f; out[34] = 0.f; out[35] = 0.f;
            if (c.rule != 0) {
#pragma unroll
                for (int o = NF; o < NO_PAD; ++o) out[o] = 0.f;
            }
            if (valid) {
                float Ai[NF], nA[NF];
                load16(A_s + (size_t)p * NF, Ai);
                const bool fire = fire_s[p] != 0;
                if (c.rule == 0) {
                    const float mult = sigmoidf_(out[2 * NF]);
#pragma unroll
                    for (int f = 0; f < NF; ++f) nA[f] = fmaf(Ai[f], sigmoidf_(out[f]), tanhf(out[NF + f]) * mult);
                } else {
#pragma unroll
                    for (int f = 0; f < NF; ++f) nA[f] = fmaf(out[f], c.orig_rate, Ai[f]);
                }
#pragma unroll
                for (int f = 0; f < NF; ++f) nA[f] = fire ? nA[f] : Ai[f];
                float4* o4 = reinterpret_cast<float4*>(An_s + (size_t)p * NF);
#pragma unroll
                for (int k = 0; k < 4; ++k) o4[k] = make_float4(nA[4 * k], nA[4 * k + 1], nA[4 * k + 2], nA[4 * k + 3]);
                float4* oo4 = reinterpret_cast<float4*>(O_s + (size_t)p * NO_PAD);
#pragma unroll
                for (int k = 0; k < NO_PAD / 4; ++k) oo4[k] = make_float4(out[4 * k], out[4 * k + 1], out[4 * k + 2], out[4 * k + 3]);
                uint4* rm4 = reinterpret_cast<uint4*>(relu_s + (size_t)p * (NH / 32));
                rm4[0] = make_uint4(rmask[0], rmask[1], rmask[2], rmask[3]);
                rm4[1] = make_uint4(rmask[4], rmask[5], rmask[6], rmask[7]);
                flags_s[p] = (uint8_t)((s.prev[it & 1][m] ? 1 : 0) | (fire ? 4 : 0));
                alive_s[p] = c.use_alpha ? (nA[3] > 0.1f ? 1 : 0) : 1;
            }
            if (tid == 0) SPHNCA_TL(15);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tbase, TMEM_COLS);
}

static int g_mlp_path = 1;
int mlp_path() { return g_mlp_path; }

int launch_step_fwd_tc(const GridView& gv, const StepConsts& c, const sphnca_step_params* p, const float* xv4,
                       const float* A_s, const uint8_t* fire_s, float* A_new_s, uint8_t* alive_s, float* Y_s,
                       float* O_s, uint8_t* flags_s, uint32_t* relu_s, cudaStream_t stream) {
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        SPHNCA_CUDA(cudaGetDevice(&dev));
        SPHNCA_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int num_tiles = (int)((gv.n + TILE - 1) / TILE);
    const int blocks = num_tiles < num_sms ? num_tiles : num_sms;
    const size_t smem = sizeof(TcSmem);
    if (gv.dim == 3) {
        SPHNCA_CUDA(cudaFuncSetAttribute(step_fwd_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SPHNCA_LAUNCH(step_fwd_tc_kernel<3>, blocks, TC_THREADS, smem, stream, gv, (const float4*)xv4, A_s, fire_s, *p, c,
                      A_new_s, Y_s, O_s, flags_s, alive_s, relu_s, num_tiles, 0);
    } else {
        SPHNCA_CUDA(cudaFuncSetAttribute(step_fwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SPHNCA_LAUNCH(step_fwd_tc_kernel<2>, blocks, TC_THREADS, smem, stream, gv, (const float4*)xv4, A_s, fire_s, *p, c,
                      A_new_s, Y_s, O_s, flags_s, alive_s, relu_s, num_tiles, 0);
    }
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

}  // namespace sphnca

extern "C" int32_t sphnca_set_mlp_path(int32_t path) {
    if (path < 0 || path > 2) { sphnca::set_error("mlp path must be 0 (CUDA cores) or 1 (tcgen05)"); return SPHNCA_ERR_INVALID; }
    sphnca::g_mlp_path = path;
    return SPHNCA_OK;
}

SPHNCA_TIMELINE_READER(sphnca_debug_timeline_fwd)
