// (tc_timing6: the forward kernel's exact MMA sequence per tile, back to back, alone and beside TMEM traffic)
// Do tcgen05.ld / tcgen05.st from the epilogue warps slow the tensor pipe down?  One warp issues the forward kernel's MMA
// mix (SS N=128 and TS N=64 / N=32, K=8 tf32) while 8 other warps run tcgen05.ld and / or tcgen05.st loops on OTHER TMEM
// columns; reported: cycles per MMA to completion and the TMEM bytes per cycle the traffic warps moved meanwhile.
#include <cstdio>
#include <cstdlib>
#include "../tc_common.cuh"
using namespace sphnca::tc;

struct Smem { float A[128 * 96]; float B[256 * 64 + 4096]; float scratch[4096]; uint64_t bar[64]; uint32_t tmem_base; volatile int stop[32]; };

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
// the kernel's order per tile: G2a(i) | G1a(i+1) | G2b(i) | G1b(i+1)
template <int VAR>
__device__ __forceinline__ void tile_mix(uint32_t tb, uint32_t a, uint32_t b, uint32_t go, int slot) {
    const uint32_t d2 = tb + 384 + 64 * slot;
    if (VAR == 2) {   // variant 2: grouped by kind: both GEMM2 halves, then both GEMM1 halves
        gemm2<0>(tb, d2, tb, b + 49152, go, true); gemm2<0>(tb, d2, tb + 128, b + 49152 + 8192, go, false);
        gemm1<0>(tb, a, b, go); gemm1<0>(tb + 128, a, b + 2048 * 6, go);
    } else {
        gemm2<VAR>(tb, d2, tb, b + 49152, go, true);
        gemm1<VAR>(tb, a, b, go);
        gemm2<VAR>(tb, d2, tb + 128, b + 49152 + 8192, go, false);
        gemm1<VAR>(tb + 128, a, b + 2048 * 6, go);
    }
}
template <int N, bool TS, int R>
__device__ __forceinline__ void issue_all(uint32_t tb, uint32_t a, uint32_t b, uint32_t go) {
    // N selects the variant here (128: kernel order, 64: N=64-only GEMM2, 32: grouped by kind); 2 tiles per call
    constexpr int VAR = N == 128 ? 0 : (N == 64 ? 1 : 2);
    tile_mix<VAR>(tb, a, b, go, 0); tile_mix<VAR>(tb, a, b, go, 1);
}

constexpr int MMA_WARP = 8, THREADS = 17 * 32, R = 96, NCFG = 1, NMODE = 8;

__global__ void __launch_bounds__(THREADS) timing_kernel(long long* out, float* gbuf) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * 96; i += THREADS) s.A[i] = 0.001f * (i % 97);
    for (int i = tid; i < 256 * 64; i += THREADS) s.B[i] = 0.002f * (i % 89);
    if (tid == 0) { for (int i = 0; i < 64; ++i) mbar_init(&s.bar[i], 1); mbar_fence_init(); for (int i = 0; i < 32; ++i) s.stop[i] = 0; }
    if (warp == MMA_WARP) tmem_alloc(&s.tmem_base, 512);
    fence_proxy_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tb = s.tmem_base;
    const uint32_t lane_addr = tb + ((uint32_t)((warp & 3) * 32) << 16);
    // mode 0: MMAs alone; 1: + ld; 2: + st; 3: + ld and st (epilogue-like: ld 16, st 16, st 16); 4: traffic alone (no MMAs, fixed count)
    for (int cfg = 0; cfg < NCFG; ++cfg) {
        for (int mode = 0; mode < NMODE; ++mode) {
            const int ph = cfg * NMODE + mode;
            __syncthreads();
            if (warp == MMA_WARP) {
                const uint32_t go = elect_one();
                const uint32_t a = smem_u32(s.A), b = smem_u32(s.B);
                long long t0 = clock64();
                if (true) {
                    for (int rep = 0; rep < 4; ++rep) {
                        if (cfg == 0) issue_all<128, false, R>(tb, a, b, go);
                        if (cfg == 1) issue_all<64, true, R>(tb, a, b, go);
                        if (cfg == 2) issue_all<32, true, R>(tb, a, b, go);
                    }
                    if (go) mma_commit(&s.bar[ph]);
                    __syncwarp();
                    mbar_wait(&s.bar[ph], 0); fence_after_sync();
                } else {
                    while (clock64() - t0 < 20000) { }
                }
                long long t1 = clock64();
                if (lane == 0) { s.stop[ph] = 1; mbar_arrive(&s.bar[40 + ph]); if (blockIdx.x == 0) out[4 * ph] = t1 - t0; }
            } else if (mode > 0 && mode < 7) {
                // helper warps: one kind of competing activity per mode, until the MMA warp is done
                long long iters = 0;
                float acc = 0.f;
                float4 v4 = make_float4(tid, 1.f, 2.f, 3.f);
                long long t0 = clock64();
                if (mode == 1) {            // 16 warps polling an mbarrier that completes only at the end
                    while (!mbar_try_wait(&s.bar[40 + ph], 0)) ++iters;
                } else if (mode == 2) {     // 16 warps: broadcast LDS.64 + FFMA loops (the epilogue's bias / gate-weight reads)
                    while (!s.stop[ph]) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) { float bx, by; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(bx), "=f"(by) : "r"(smem_u32(&s.scratch[(2 * j + 2 * (int)iters) & 1022]))); acc = fmaf(bx, by, acc); }
                        ++iters;
                    }
                } else if (mode == 3) {     // 16 warps: STS.128 streams (the loader's tile writes)
                    while (!s.stop[ph]) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(smem_u32(&s.scratch[((j * 32 + lane) * 4) & 4095])), "f"(v4.x), "f"(v4.y), "f"(v4.z), "f"(v4.w) : "memory");
                        ++iters;
                    }
                } else if (mode == 4) {     // 16 warps: row-strided 16-byte global stores (the finalizer's rows)
                    float4* g = reinterpret_cast<float4*>(gbuf) + (size_t)(blockIdx.x * THREADS + tid) * 64;
                    while (!s.stop[ph]) {
#pragma unroll
                        for (int j = 0; j < 13; ++j) g[(j + iters) & 63] = v4;
                        ++iters;
                    }
                } else if (mode == 5) {     // 16 warps: global loads of 192-byte rows (the loader's reads)
                    const float4* g = reinterpret_cast<const float4*>(gbuf) + (size_t)(blockIdx.x * THREADS + tid) * 64;
                    while (!s.stop[ph]) {
#pragma unroll
                        for (int j = 0; j < 12; ++j) { const float4 q = __ldcg(g + ((j + iters) & 63)); acc += q.x; }
                        ++iters;
                    }
                } else {                    // 6: 8 warps the epilogue's TMEM pattern + 8 warps polling
                    if (warp < 8) {
                        uint32_t r[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) r[j] = tid + j;
                        const uint32_t col = 384 + 64 * (warp >> 2);
                        while (!s.stop[ph]) { tmem_ld16(lane_addr + col, r); wait_ld(); r[3] += 1; tmem_st16(lane_addr + col + 16, r); tmem_st16(lane_addr + col + 32, r); wait_st(); ++iters; }
                        if (r[5] == 0x7fffffff) acc += r[0];
                    } else {
                        while (!mbar_try_wait(&s.bar[40 + ph], 0)) ++iters;
                    }
                }
                long long t1 = clock64();
                if (lane == 0 && warp == 0 && blockIdx.x == 0) { out[4 * ph + 1] = iters; out[4 * ph + 2] = t1 - t0; }
                if (acc == 1.2345f) out[63] = (long long)acc;
            }
        }
    }
    fence_before_sync(); __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tb, 512);
}

int main(int argc, char** argv) {
    const int grid = argc > 1 ? atoi(argv[1]) : 1;
    long long* d; cudaMalloc(&d, 256 * 8); cudaMemset(d, 0, 256 * 8);
    float* gbuf; cudaMalloc(&gbuf, (size_t)grid * THREADS * 64 * 16); cudaMemset(gbuf, 0, (size_t)grid * THREADS * 64 * 16);
    cudaFuncSetAttribute(timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    timing_kernel<<<grid, THREADS, sizeof(Smem)>>>(d, gbuf);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[256]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* nm[NCFG] = {"kernel order"};
    const char* md[NMODE] = {"alone", "+ 16 warps polling", "+ 16 warps LDS+FFMA", "+ 16 warps STS.128", "+ 16 warps STG rows", "+ 16 warps LDG rows", "+ 8 TMEM + 8 polling", "alone again"};
    // bytes per traffic iteration per warp: mode 1: 2 x 16 cols x 32 lanes x 4 B; mode 2: same; mode 3: 3 x
    
    for (int c = 0; c < NCFG; ++c)
        for (int m = 0; m < NMODE; ++m) {
            const int ph = c * NMODE + m;
            printf("%-18s %-22s: %7.1f cyc/tile to completion", nm[c], md[m], h[4 * ph] / 8.0);
            if (m > 0 && m < 7) printf("   helpers: %lld iters/warp in %lld cyc", h[4 * ph + 1], h[4 * ph + 2]);
            printf("\n");
        }
    return 0;
}
