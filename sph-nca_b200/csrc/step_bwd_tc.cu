// Data gradient of the NCA step after the perception sweep, on the Blackwell tensor path.
// Same maths as step_bwd_kernel (step.cu); no hidden-layer recompute: the forward saved the ReLU
// activity bits (32 B / particle).
//
//   loader warps (4)   thread = particle: gate / tanh / fire / life-mask derivatives -> dO[36] (written to
//                      HBM for the weight-gradient stage and, tf32 hi/lo split, into the MMA A tile),
//                      the direct part of dL/dA, the scalar-gate column dO[32], the ReLU bits -> smem.
//   MMA warp (1 lane)  GEMM3  D3[128p x 128h] = dO[128 x 32] . W2t_half^T      (B = W2^T stored K-major)
//                      GEMM4  Dy[128p x 48]  += dH[128 x 128](TMEM) . W1t_half^T (B = W1^T stored K-major)
//   epilogue warps (4) tcgen05.ld D3 -> (+ dO[32]*W2[32][h]) -> ReLU bits -> tf32 hi/lo -> tcgen05.st dH
//                      IN PLACE over D3; final: tcgen05.ld Dy -> dL/dY (32) and dL/dA direct part (16).
//
// (MN-major tf32 operands need the SW128_32B layout, so instead of re-viewing the forward's W1/W2 tiles
//  the kernel stages transposed K-major copies: W2^T 64 KB + W1^T 96 KB, hi/lo.)
// GEMM4 has only 48 output columns, where an MMA costs ~32 cycles + N/4 whatever N is (probes/tc_timing4: A from
// TMEM, N=48 59 cycles, N=128 71), so the hi and lo copies of W1^T are STACKED along N (96 rows): dH_hi . [W_hi ; W_lo]
// is one MMA into Dy columns 0..47 | 48..95 and dH_lo . W_hi adds to columns 0..47 -- 2 MMAs per k-step instead of 3;
// the final epilogue adds the two column groups.
// TMEM: D3/dH_hi 128 + dH_lo 128 + Dy 96 = 352 columns.  MMAs execute in issue order, so GEMM3 of the
// next half may be issued right behind GEMM4 of the previous one on the same columns.
#include "step_common.cuh"
#include "tc_common.cuh"

namespace sphnca {

using namespace tc;

namespace bw {
constexpr int TILE = 128;
// epilogue warps 0..7: warp & 3 = TMEM lane quarter, warp >> 2 = which 64 of a half's 128 hidden columns
constexpr int EPI_WARPS = 8, MMA_WARP = 8, LOAD_WARP0 = 9, LOAD_WARPS = 4;
constexpr int THREADS = (LOAD_WARP0 + LOAD_WARPS) * 32;      // 416
constexpr int KO = 32;                                       // outputs handled by the tensor core
constexpr uint32_t COL_D3 = 0, COL_DHLO = 128, COL_DY = 256, TMEM_COLS = 512;
enum { DO_FULL = 0, DO_EMPTY, D3_FULL, DH_FULL, DY_FULL, DY_EMPTY, NUM_BARS };

struct Smem {
    float W2t_hi[2][128 * KO], W2t_lo[2][128 * KO];       // [hidden half] rows = h (128), K = o (32)
    float W1t[2][2 * NIN * 128];                          // [hidden half] rows 0..47 = k hi, 48..95 = k lo ; K = h (128)
    float dO_hi[TILE * KO], dO_lo[TILE * KO];             // rows = particle, K = o
    float dAdir[2][TILE][NF];
    float dO32[2][TILE];
    uint32_t relu[2][TILE][NH / 32];
    float w2_last[NH];
    uint64_t bar[NUM_BARS];
    uint32_t tmem_base;
};
__device__ __forceinline__ uint32_t fidx(uint32_t rows, uint32_t m, uint32_t k) { return chunk_off(rows, m, k >> 2) / 4 + (k & 3); }
}  // namespace bw

__global__ void __launch_bounds__(bw::THREADS, 1)
step_bwd_tc_kernel(int n_total, const float* __restrict__ A_s, const float* __restrict__ O_s,
                   const uint8_t* __restrict__ flags_s, const uint32_t* __restrict__ relu_s,
                   const float* __restrict__ dAn_s, sphnca_step_params prm, StepConsts c,
                   float* __restrict__ dO_s, float* __restrict__ dY_s, float* __restrict__ dAdir_s, int num_tiles) {
    using namespace bw;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int n_mma = c.n_out < KO ? c.n_out : KO;
    for (int i = tid; i < KO * NH; i += THREADS) {            // W2 [o][h] -> W2t[half] (rows h, K = o)
        const int o_ = i / NH, h = i - o_ * NH;
        float hi, lo; split_tf32(o_ < n_mma ? __ldg(prm.W2 + (size_t)o_ * NH + h) : 0.f, hi, lo);
        const uint32_t o = fidx(128, h & 127, o_);
        s.W2t_hi[h >> 7][o] = hi; s.W2t_lo[h >> 7][o] = lo;
    }
    for (int i = tid; i < NH * NIN; i += THREADS) {           // W1 [h][k] -> W1t[half] (rows k, K = h)
        const int h = i / NIN, k = i - h * NIN;
        float hi, lo; split_tf32(__ldg(prm.W1 + i), hi, lo);
        s.W1t[h >> 7][fidx(2 * NIN, k, h & 127)] = hi; s.W1t[h >> 7][fidx(2 * NIN, NIN + k, h & 127)] = lo;
    }
    for (int i = tid; i < NH; i += THREADS) s.w2_last[i] = c.n_out > KO ? __ldg(prm.W2 + (size_t)KO * NH + i) : 0.f;
    if (tid == 0) {
        mbar_init(&s.bar[DO_FULL], LOAD_WARPS * 32); mbar_init(&s.bar[DO_EMPTY], 1);
        mbar_init(&s.bar[D3_FULL], 1);               mbar_init(&s.bar[DH_FULL], EPI_WARPS * 32);
        mbar_init(&s.bar[DY_FULL], 1);               mbar_init(&s.bar[DY_EMPTY], 4 * 32);
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(&s.tmem_base, TMEM_COLS);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = s.tmem_base;

    if (warp >= LOAD_WARP0) {
        // ======================= prologue: dL/d(MLP output) and the direct dL/dA part =======================
        const int m = tid - LOAD_WARP0 * 32;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int p = tile * TILE + m;
            const bool valid = p < n_total;
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
                load16(A_s + (size_t)p * NF, Ai);
                load16(dAn_s + (size_t)p * NF, dnA);
                const uint8_t fl = flags_s[p];
                const bool living = (fl & 1) && (fl & 2), fire = fl & 4;
#pragma unroll
                for (int f = 0; f < NF; ++f) dnA[f] = living ? dnA[f] : 0.f;
                if (fire) {
                    const float4* o4 = reinterpret_cast<const float4*>(O_s + (size_t)p * NO_PAD);
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
                    const uint4* r4 = reinterpret_cast<const uint4*>(relu_s + (size_t)p * (NH / 32));
                    const uint4 r0 = __ldg(r4), r1 = __ldg(r4 + 1);
                    rbits[0] = r0.x; rbits[1] = r0.y; rbits[2] = r0.z; rbits[3] = r0.w;
                    rbits[4] = r1.x; rbits[5] = r1.y; rbits[6] = r1.z; rbits[7] = r1.w;
                } else {
#pragma unroll
                    for (int f = 0; f < NF; ++f) dAdir[f] = dnA[f];
                }
                float4* do4 = reinterpret_cast<float4*>(dO_s + (size_t)p * NO_PAD);
#pragma unroll
                for (int k = 0; k < NO_PAD / 4; ++k) do4[k] = make_float4(dO[4*k], dO[4*k+1], dO[4*k+2], dO[4*k+3]);
            }
            mbar_wait(&s.bar[DO_EMPTY], (it & 1) ^ 1);        // GEMM3 of the previous tile consumed the dO tile
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
            const uint32_t ah = smem_u32(s.dO_hi), al = smem_u32(s.dO_lo);
            auto gemm3 = [&](int half) {
                const uint32_t wh = smem_u32(s.W2t_hi[half]), wl = smem_u32(s.W2t_lo[half]);
#pragma unroll
                for (int ks = 0; ks < KO / 8; ++ks) {
                    const uint64_t a_hi = make_smem_desc(ah + chunk_off(TILE, 0, 2 * ks), TILE * 16, 128);
                    const uint64_t a_lo = make_smem_desc(al + chunk_off(TILE, 0, 2 * ks), TILE * 16, 128);
                    const uint64_t b_hi = make_smem_desc(wh + chunk_off(128, 0, 2 * ks), 128 * 16, 128);
                    const uint64_t b_lo = make_smem_desc(wl + chunk_off(128, 0, 2 * ks), 128 * 16, 128);
                    mma_tf32_ss_e(tbase + COL_D3, a_lo, b_hi, idesc3, ks > 0, go);
                    mma_tf32_ss_e(tbase + COL_D3, a_hi, b_lo, idesc3, true, go);
                    mma_tf32_ss_e(tbase + COL_D3, a_hi, b_hi, idesc3, true, go);
                }
            };
            auto gemm4 = [&](int half) {
                const uint32_t w = smem_u32(s.W1t[half]);
#pragma unroll
                for (int ks = 0; ks < 128 / 8; ++ks) {
                    // rows 0..47 of this tile = W1^T hi, all 96 rows = the stacked [hi ; lo] tile
                    const uint64_t b = make_smem_desc(w + chunk_off(2 * NIN, 0, 2 * ks), 2 * NIN * 16, 128);
                    mma_tf32_ts_e(tbase + COL_DY, tbase + COL_D3 + 8 * ks, b, idesc4c, (half | ks) > 0, go);   // dH_hi . [W_hi ; W_lo]
                    mma_tf32_ts_e(tbase + COL_DY, tbase + COL_DHLO + 8 * ks, b, idesc4, true, go);             // dH_lo . W_hi
                }
            };
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                mbar_wait(&s.bar[DO_FULL], it & 1);
                fence_after_sync();
                gemm3(0);
                mma_commit_e(&s.bar[D3_FULL], go);
                mbar_wait(&s.bar[DH_FULL], 0);
                fence_after_sync();
                mbar_wait(&s.bar[DY_EMPTY], (it & 1) ^ 1);    // epilogue drained Dy of the previous tile
                fence_after_sync();
                gemm4(0);
                gemm3(1);                                     // in issue order behind GEMM4(0) on the same columns
                mma_commit_e(&s.bar[D3_FULL], go);
                mma_commit_e(&s.bar[DO_EMPTY], go);
                mbar_wait(&s.bar[DH_FULL], 1);
                fence_after_sync();
                gemm4(1);
                mma_commit_e(&s.bar[DY_FULL], go);
            }
        }
    } else {
        // ======================= epilogue =======================
        const int m = (warp & 3) * 32 + lane;
        const int chalf = warp >> 2;
        const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int p = tile * TILE + m;
            const bool valid = p < n_total;
            const int slot = it & 1;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                mbar_wait(&s.bar[D3_FULL], half);
                fence_after_sync();
                const float d32 = s.dO32[slot][m];
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int cc = chalf * 64 + c4 * 16;
                    uint32_t r[16], hi[16], lo[16];
                    tmem_ld16(lane_addr + COL_D3 + cc, r);
                    wait_ld();
                    const uint32_t bits = s.relu[slot][m][(half * 128 + cc) >> 5] >> (cc & 16);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = half * 128 + cc + j;
                        const float dh = ((bits >> j) & 1u) ? fmaf(d32, s.w2_last[n], __uint_as_float(r[j])) : 0.f;
                        float h_, l_; split_tf32(dh, h_, l_);
                        hi[j] = __float_as_uint(h_); lo[j] = __float_as_uint(l_);
                    }
                    tmem_st16(lane_addr + COL_D3 + cc, hi);       // in place: D3 -> dH_hi
                    tmem_st16(lane_addr + COL_DHLO + cc, lo);
                }
                wait_st();
                fence_before_sync();
                mbar_arrive(&s.bar[DH_FULL]);
            }
            if (chalf != 0) continue;                         // the final part belongs to the first column-half warps
            mbar_wait(&s.bar[DY_FULL], it & 1);
            fence_after_sync();
            float dy[NIN];
#pragma unroll
            for (int cc = 0; cc < NIN; cc += 16) {
                uint32_t r[16], r2[16];
                tmem_ld16(lane_addr + COL_DY + cc, r); tmem_ld16(lane_addr + COL_DY + NIN + cc, r2); wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) dy[cc + j] = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
            }
            fence_before_sync();
            mbar_arrive(&s.bar[DY_EMPTY]);
            if (valid) {
                float4* d4 = reinterpret_cast<float4*>(dAdir_s + (size_t)p * NF);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 a = *reinterpret_cast<const float4*>(&s.dAdir[slot][m][4 * k]);
                    d4[k] = make_float4(a.x + dy[4*k], a.y + dy[4*k+1], a.z + dy[4*k+2], a.w + dy[4*k+3]);
                }
                float4* g4 = reinterpret_cast<float4*>(dY_s + (size_t)p * 2 * NF);
#pragma unroll
                for (int k = 0; k < 8; ++k) g4[k] = make_float4(dy[NF+4*k], dy[NF+4*k+1], dy[NF+4*k+2], dy[NF+4*k+3]);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tbase, bw::TMEM_COLS);
}

int launch_step_bwd_tc(int64_t N, const StepConsts& c, const sphnca_step_params* p, const float* A_s, const float* O_s,
                       const uint8_t* flags_s, const uint32_t* relu_s, const float* dA_next_s, float* dO_s, float* dY_s,
                       float* dAdir_s, cudaStream_t stream) {
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
                  dO_s, dY_s, dAdir_s, num_tiles);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

}  // namespace sphnca
