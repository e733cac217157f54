// (tc_timing8: when do tcgen05.commit arrivals placed BETWEEN the MMA groups of one tile fire?)
// Do tcgen05.ld / tcgen05.st from the epilogue warps slow the tensor pipe down?  One warp issues the forward kernel's MMA
// mix (SS N=128 and TS N=64 / N=32, K=8 tf32) while 8 other warps run tcgen05.ld and / or tcgen05.st loops on OTHER TMEM
// columns; reported: cycles per MMA to completion and the TMEM bytes per cycle the traffic warps moved meanwhile.
#include <cstdio>
#include <cstdlib>
#include "../tc_common.cuh"
using namespace sphnca::tc;

struct Smem { float A[128 * 96]; float B[256 * 64 + 16384]; uint64_t bar[64]; uint32_t tmem_base; volatile int stop[32]; };

__device__ __forceinline__ void mma_ss_p(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t go) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, 1;\n\t}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(go) : "memory");
}
__device__ __forceinline__ void mma_ts_p(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t go) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, 1;\n\t}\n"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(go) : "memory");
}
__device__ __forceinline__ void mma_ss_a(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc, uint32_t go) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(go) : "memory");
}
// one hidden half of GEMM1: 6 k-steps x 3 products, SS N=128 -> D1
template <int VAR>
__device__ __forceinline__ void gemm1(uint32_t d1, uint32_t a, uint32_t b, uint32_t go) {
    constexpr uint32_t idesc = make_idesc_tf32(128, 128);
#pragma unroll
    for (int ks = 0; ks < 6; ++ks) {
        const uint64_t ad = make_smem_desc(a + chunk_off(128, 0, 2 * ks), 2048, 128), bd = make_smem_desc(b + chunk_off(128, 0, 2 * ks), 2048, 128);
        const uint64_t ad2 = make_smem_desc(a + 24576 + chunk_off(128, 0, 2 * ks), 2048, 128), bd2 = make_smem_desc(b + 24576 + chunk_off(128, 0, 2 * ks), 2048, 128);
        mma_ss_a(d1, ad2, bd, idesc, ks > 0, go);
        mma_ss_a(d1, ad, bd2, idesc, 1, go);
        mma_ss_a(d1, ad, bd, idesc, 1, go);
    }
}
// one hidden half of GEMM2: 16 x TS N=32 (H_lo) -> d2+32, then 16 x TS N=64 (H_hi) -> d2
template <int VAR>
__device__ __forceinline__ void gemm2(uint32_t tb, uint32_t d2, uint32_t hhi, uint32_t b, uint32_t go, bool first) {
    constexpr uint32_t i32 = make_idesc_tf32(128, 32), i64 = make_idesc_tf32(128, 64);
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
        const uint64_t bd = make_smem_desc(b + chunk_off(64, 0, 2 * ks), 64 * 16, 128);
        if (VAR == 1) mma_ts_p(d2, tb + 256 + 8 * ks, bd, i64, go);          // variant 1: every GEMM2 MMA N=64 into the same accumulator
        else if (first && ks == 0) { asm volatile("" ::: "memory"); mma_ts_p(d2, tb + 256, bd, i64, go); }
        else mma_ts_p(d2 + 32, tb + 256 + 8 * ks, bd, i32, go);
    }
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
        const uint64_t bd = make_smem_desc(b + chunk_off(64, 0, 2 * ks), 64 * 16, 128);
        mma_ts_p(d2, hhi + 8 * ks, bd, i64, go);
    }
}

constexpr int MMA_WARP = 8, THREADS = 9 * 32, NB = 6;

__global__ void __launch_bounds__(THREADS) timing_kernel(long long* out, int reps) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * 96; i += THREADS) s.A[i] = 0.001f * (i % 97);
    for (int i = tid; i < 256 * 64; i += THREADS) s.B[i] = 0.002f * (i % 89);
    if (tid == 0) { for (int i = 0; i < 64; ++i) mbar_init(&s.bar[i], 1); mbar_fence_init(); }
    if (warp == MMA_WARP) tmem_alloc(&s.tmem_base, 512);
    fence_proxy_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tb = s.tmem_base;
    __shared__ long long t_start;
    for (int rep = 0; rep < reps; ++rep) {
        __syncthreads();
        if (warp == MMA_WARP) {
            const uint32_t go = elect_one();
            const uint32_t a = smem_u32(s.A), b = smem_u32(s.B);
            const uint32_t d2 = tb + 384;
            constexpr uint32_t i32 = make_idesc_tf32(128, 32), i64 = make_idesc_tf32(128, 64);
            const long long t0 = clock64();
            if (lane == 0) t_start = t0;
            // G2a: H_lo part | commit 0 | H_hi part
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const uint32_t w = b + 49152 + 8192 * half, hhi = tb + 128 * half;
#pragma unroll
                for (int ks = 0; ks < 16; ++ks) {
                    const uint64_t bd = make_smem_desc(w + chunk_off(64, 0, 2 * ks), 64 * 16, 128);
                    if (half == 0 && ks == 0) mma_ts_p(d2, tb + 256, bd, i64, go);
                    else mma_ts_p(d2 + 32, tb + 256 + 8 * ks, bd, i32, go);
                }
                if (go) mma_commit(&s.bar[3 * half + 0]);
#pragma unroll
                for (int ks = 0; ks < 16; ++ks) {
                    const uint64_t bd = make_smem_desc(w + chunk_off(64, 0, 2 * ks), 64 * 16, 128);
                    mma_ts_p(d2, hhi + 8 * ks, bd, i64, go);
                }
                if (half == 1 && go) mma_commit(&s.bar[3 * half + 1]);
                gemm1<0>(tb + 128 * half, a, b + 12288 * half, go);
                if (go) mma_commit(&s.bar[3 * half + (half ? 2 : 1)]);
                if (half == 0 && go) mma_commit(&s.bar[2]);       // (a second commit at the same point)
            }
            const long long t1 = clock64();
            if (lane == 0 && blockIdx.x == 0) out[rep * 8 + 7] = t1 - t0;
        } else if (warp < NB) {
            mbar_wait(&s.bar[warp], rep & 1);
            const long long t = clock64();
            if (lane == 0 && blockIdx.x == 0) out[rep * 8 + warp] = t - *(volatile long long*)&t_start;
        }
        __syncthreads();
    }
    fence_before_sync(); __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tb, 512);
}

int main(int argc, char** argv) {
    const int grid = argc > 1 ? atoi(argv[1]) : 1, reps = 4;
    long long* d; cudaMalloc(&d, 256 * 8); cudaMemset(d, 0, 256 * 8);
    cudaFuncSetAttribute(timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    timing_kernel<<<grid, THREADS, sizeof(Smem)>>>(d, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[256]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* nm[NB] = {"G2a H_lo (16 x N32)", "+ G2a H_hi (16 x N64) + G1a (18 x N128)", "same point, 2nd commit", "+ G2b H_lo", "+ G2b H_hi", "+ G1b"};
    const int expect[NB] = {290, 1980, 1980, 2270, 2790, 3960};
    for (int rep = 0; rep < reps; ++rep) {
        printf("rep %d: all issued after %lld cycles\n", rep, h[rep * 8 + 7]);
        for (int k = 0; k < NB; ++k) printf("    commit %d fired at %6lld (tensor-pipe floor ~%d)   %s\n", k, h[rep * 8 + k], expect[k], nm[k]);
    }
    return 0;
}
