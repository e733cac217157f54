// Per-instruction cost of the tcgen05.mma shapes used by the fused MLP (B200 micro-benchmark).
// One CTA, one issuing thread: R back-to-back MMAs into the same accumulator, one commit, wait.
#include <cstdio>
#include <cstdlib>
#include "../tc_common.cuh"
using namespace sphnca::tc;

struct Smem { float A[128 * 64]; float B[256 * 64]; uint64_t bar[16]; uint32_t tmem_base; };

__device__ __forceinline__ void mma_bf16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, bool acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)acc) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t M, uint32_t N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24); }

__global__ void __launch_bounds__(128) timing_kernel(long long* out, int R) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 64; i += 128) s.A[i] = 0.001f * (i % 97);
    for (int i = tid; i < 256 * 64; i += 128) s.B[i] = 0.002f * (i % 89);
    if (tid == 0) { for (int i = 0; i < 16; ++i) mbar_init(&s.bar[i], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&s.tmem_base, 512);
    fence_proxy_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tb = s.tmem_base;
    if (tid == 0) {
        const uint32_t a = smem_u32(s.A), b = smem_u32(s.B);
        int cfg = 0;
        auto run = [&](auto issue) {
            // warm-up
            issue(0); mma_commit(&s.bar[cfg]); mbar_wait(&s.bar[cfg], 0); fence_after_sync();
            long long t0 = clock64();
            for (int r = 0; r < R; ++r) issue(r);
            mma_commit(&s.bar[cfg]); long long t1 = clock64();
            mbar_wait(&s.bar[cfg], 1); fence_after_sync();
            long long t2 = clock64();
            out[2 * cfg] = t1 - t0; out[2 * cfg + 1] = t2 - t0; ++cfg;
        };
        // 0: tf32 SS M128 N128 K8   1: tf32 SS N256   2: tf32 SS N32   3: tf32 TS N32   4: tf32 TS N128
        // 5: bf16 SS N128 K16       6: bf16 SS N256   7: tf32 SS N64   8: tf32 TS N256
        run([&](int r) { mma_tf32_ss(tb, make_smem_desc(a + chunk_off(128, 0, 2 * (r % 6)), 2048, 128), make_smem_desc(b + chunk_off(128, 0, 2 * (r % 6)), 2048, 128), make_idesc_tf32(128, 128), true); });
        run([&](int r) { mma_tf32_ss(tb, make_smem_desc(a + chunk_off(128, 0, 2 * (r % 6)), 2048, 128), make_smem_desc(b + chunk_off(256, 0, 2 * (r % 6)), 4096, 128), make_idesc_tf32(128, 256), true); });
        run([&](int r) { mma_tf32_ss(tb, make_smem_desc(a + chunk_off(128, 0, 2 * (r % 6)), 2048, 128), make_smem_desc(b + chunk_off(32, 0, 2 * (r % 16)), 512, 128), make_idesc_tf32(128, 32), true); });
        run([&](int r) { mma_tf32_ts(tb + 384, tb + 8 * (r % 16), make_smem_desc(b + chunk_off(32, 0, 2 * (r % 16)), 512, 128), make_idesc_tf32(128, 32), true); });
        run([&](int r) { mma_tf32_ts(tb + 384, tb + 8 * (r % 16), make_smem_desc(b + chunk_off(128, 0, 2 * (r % 6)), 2048, 128), make_idesc_tf32(128, 128), true); });
        run([&](int r) { mma_bf16_ss(tb, make_smem_desc(a + chunk_off(128, 0, 2 * (r % 6)), 2048, 128), make_smem_desc(b + chunk_off(128, 0, 2 * (r % 6)), 2048, 128), idesc_bf16(128, 128), true); });
        run([&](int r) { mma_bf16_ss(tb, make_smem_desc(a + chunk_off(128, 0, 2 * (r % 6)), 2048, 128), make_smem_desc(b + chunk_off(256, 0, 2 * (r % 6)), 4096, 128), idesc_bf16(128, 256), true); });
        run([&](int r) { mma_tf32_ss(tb, make_smem_desc(a + chunk_off(128, 0, 2 * (r % 6)), 2048, 128), make_smem_desc(b + chunk_off(64, 0, 2 * (r % 6)), 1024, 128), make_idesc_tf32(128, 64), true); });
        run([&](int r) { mma_tf32_ts(tb + 256, tb + 8 * (r % 16), make_smem_desc(b + chunk_off(256, 0, 2 * (r % 6)), 4096, 128), make_idesc_tf32(128, 256), true); });
    }
    fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
    const char* names[] = {"tf32 SS M128 N128 K8", "tf32 SS M128 N256 K8", "tf32 SS M128 N32  K8", "tf32 TS M128 N32  K8", "tf32 TS M128 N128 K8",
                           "bf16 SS M128 N128 K16", "bf16 SS M128 N256 K16", "tf32 SS M128 N64  K8", "tf32 TS M128 N256 K8"};
    long long* d; cudaMalloc(&d, 64 * 8);
    cudaFuncSetAttribute(timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    for (int R : {32, 128}) {
        timing_kernel<<<1, 128, sizeof(Smem)>>>(d, R);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("R = %d back-to-back MMAs\n", R);
        for (int i = 0; i < 9; ++i) printf("  %-24s issue %6lld cyc (%5.1f/MMA)   complete %6lld cyc (%5.1f/MMA)\n", names[i], h[2 * i], (double)h[2 * i] / R, h[2 * i + 1], (double)h[2 * i + 1] / R);
    }
    return 0;
}
