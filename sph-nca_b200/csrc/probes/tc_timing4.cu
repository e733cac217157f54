// Is the ~72-90 cycle per-MMA floor of tc_timing.cu / tc_timing2.cu the issuing THREAD (descriptor arithmetic in a
// rolled loop) rather than the tensor core?  Same MMAs, fully unrolled with precomputed descriptors.
#include <cstdio>
#include <cstdlib>
#include "../tc_common.cuh"
using namespace sphnca::tc;

struct Smem { float A[128 * 64]; float B[256 * 64]; uint64_t bar[64]; uint32_t tmem_base; };

// (elect_one: tc_common.cuh)
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
        if (TS) mma_ts_p(tb + 256, tb + 8 * (r % 16), bdr, idesc, go);
        else mma_ss_p(tb, ad + (uint64_t)(chunk_off(128, 0, 2 * (r % 6)) >> 4), bdr, idesc, go);
    }
}

__global__ void __launch_bounds__(128) timing_kernel(long long* out) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 64; i += 128) s.A[i] = 0.001f * (i % 97);
    for (int i = tid; i < 256 * 64; i += 128) s.B[i] = 0.002f * (i % 89);
    if (tid == 0) { for (int i = 0; i < 64; ++i) mbar_init(&s.bar[i], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&s.tmem_base, 512);
    fence_proxy_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tb = s.tmem_base;
    if (warp == 0) {
        const uint32_t go = elect_one();
        const uint32_t a = smem_u32(s.A), b = smem_u32(s.B);
        int cfg = 0;
        auto run = [&](auto issue) {
            issue(); if (go) mma_commit(&s.bar[cfg]); __syncwarp(); mbar_wait(&s.bar[cfg], 0); fence_after_sync();
            long long t0 = clock64();
            issue();
            long long t1 = clock64();
            if (go) mma_commit(&s.bar[cfg]);
            __syncwarp();
            mbar_wait(&s.bar[cfg], 1); fence_after_sync();
            long long t2 = clock64();
            if (go) { out[2 * cfg] = t1 - t0; out[2 * cfg + 1] = t2 - t0; } ++cfg;
        };
        constexpr int R = 96;
        run([&] { issue_all<32, false, R>(tb, a, b, go); });
        run([&] { issue_all<48, false, R>(tb, a, b, go); });
        run([&] { issue_all<64, false, R>(tb, a, b, go); });
        run([&] { issue_all<128, false, R>(tb, a, b, go); });
        run([&] { issue_all<256, false, R>(tb, a, b, go); });
        run([&] { issue_all<32, true, R>(tb, a, b, go); });
        run([&] { issue_all<48, true, R>(tb, a, b, go); });
        run([&] { issue_all<64, true, R>(tb, a, b, go); });
        run([&] { issue_all<128, true, R>(tb, a, b, go); });
        run([&] { issue_all<256, true, R>(tb, a, b, go); });
    }
    fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 64 * 8); cudaMemset(d, 0, 64 * 8);
    cudaFuncSetAttribute(timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    timing_kernel<<<1, 128, sizeof(Smem)>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* nm[10] = {"SS N32", "SS N48", "SS N64", "SS N128", "SS N256", "TS N32", "TS N48", "TS N64", "TS N128", "TS N256"};
    for (int c = 0; c < 10; ++c)
        printf("tf32 %-8s M128 K8 unrolled x96: issue %6.1f cyc/MMA   complete %6.1f cyc/MMA\n", nm[c], h[2 * c] / 96.0, h[2 * c + 1] / 96.0);
    return 0;
}
