// Do tcgen05.ld / tcgen05.st from the epilogue warps slow the tensor pipe down?  One warp issues the forward kernel's MMA
// mix (SS N=128 and TS N=64 / N=32, K=8 tf32) while 8 other warps run tcgen05.ld and / or tcgen05.st loops on OTHER TMEM
// columns; reported: cycles per MMA to completion and the TMEM bytes per cycle the traffic warps moved meanwhile.
#include <cstdio>
#include <cstdlib>
#include "../tc_common.cuh"
using namespace sphnca::tc;

struct Smem { float A[128 * 64]; float B[256 * 64]; uint64_t bar[64]; uint32_t tmem_base; volatile int stop[32]; };

__device__ __forceinline__ void mma_ss_p(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t go) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, 1;\n\t}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(go) : "memory");
}
__device__ __forceinline__ void mma_ts_p(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t go) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, 1;\n\t}\n"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(go) : "memory");
}
template <int N, bool TS, int R>
__device__ __forceinline__ void issue_all(uint32_t tb, uint32_t a, uint32_t b, uint32_t go) {
    const uint64_t ad = make_smem_desc(a, 2048, 128), bd = make_smem_desc(b, N * 16, 128);
    constexpr uint32_t idesc = make_idesc_tf32(128, N);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint64_t bdr = bd + (uint64_t)(chunk_off(N, 0, 2 * (r % 6)) >> 4);
        if (TS) mma_ts_p(tb + 256, tb + 128 + 8 * (r % 16), bdr, idesc, go);
        else mma_ss_p(tb, ad + (uint64_t)(chunk_off(128, 0, 2 * (r % 6)) >> 4), bdr, idesc, go);
    }
}

constexpr int TRAFFIC_WARPS = 8, MMA_WARP = 8, THREADS = 9 * 32, R = 96, NCFG = 3, NMODE = 5;

__global__ void __launch_bounds__(THREADS) timing_kernel(long long* out) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * 64; i += THREADS) s.A[i] = 0.001f * (i % 97);
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
                if (mode < 4) {
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
                if (lane == 0) { s.stop[ph] = 1; out[4 * ph] = t1 - t0; }
            } else if (mode > 0) {
                uint32_t r[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = tid + j;
                long long iters = 0;
                const uint32_t col = 384 + 64 * (warp >> 2);
                long long t0 = clock64();
                while (!s.stop[ph]) {
                    if (mode == 1) { tmem_ld16(lane_addr + col, r); wait_ld(); tmem_ld16(lane_addr + col + 16, r); wait_ld(); }
                    else if (mode == 2) { tmem_st16(lane_addr + col, r); tmem_st16(lane_addr + col + 16, r); wait_st(); }
                    else { tmem_ld16(lane_addr + col, r); wait_ld(); r[3] += 1; tmem_st16(lane_addr + col + 16, r); tmem_st16(lane_addr + col + 32, r); wait_st(); }
                    ++iters;
                }
                long long t1 = clock64();
                if (lane == 0 && warp == 0) { out[4 * ph + 1] = iters; out[4 * ph + 2] = t1 - t0; }
                if (r[5] == 0x7fffffff) out[63] = r[0];
            }
        }
    }
    fence_before_sync(); __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tb, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 256 * 8); cudaMemset(d, 0, 256 * 8);
    cudaFuncSetAttribute(timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    timing_kernel<<<1, THREADS, sizeof(Smem)>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[256]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* nm[NCFG] = {"SS N128", "TS N64", "TS N32"};
    const char* md[NMODE] = {"alone", "+ 8 warps ld", "+ 8 warps st", "+ 8 warps ld,st,st", "traffic alone"};
    // bytes per traffic iteration per warp: mode 1: 2 x 16 cols x 32 lanes x 4 B; mode 2: same; mode 3: 3 x
    const double bytes_iter[NMODE] = {0, 2 * 2048.0, 2 * 2048.0, 3 * 2048.0, 3 * 2048.0};
    for (int c = 0; c < NCFG; ++c)
        for (int m = 0; m < NMODE; ++m) {
            const int ph = c * NMODE + m;
            printf("tf32 %-8s %-22s: %6.1f cyc/MMA to completion", nm[c], md[m], m < 4 ? h[4 * ph] / (4.0 * R) : 0.0);
            if (m > 0) printf("   traffic %6.1f B/cyc over 8 warps (%lld iters/warp in %lld cyc)", 8 * bytes_iter[m] * h[4 * ph + 1] / (double)h[4 * ph + 2], h[4 * ph + 1], h[4 * ph + 2]);
            printf("\n");
        }
    return 0;
}
