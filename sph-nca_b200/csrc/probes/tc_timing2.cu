// Does the per-MMA floor of small-N tcgen05.mma kind::tf32 come from the dependency on the accumulator?
// R back-to-back MMAs rotating over G independent accumulators (distinct TMEM column blocks).
#include <cstdio>
#include <cstdlib>
#include "../tc_common.cuh"
using namespace sphnca::tc;

struct Smem { float A[128 * 64]; float B[256 * 64]; uint64_t bar[64]; uint32_t tmem_base; };

__global__ void __launch_bounds__(128) timing_kernel(long long* out, int R) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 64; i += 128) s.A[i] = 0.001f * (i % 97);
    for (int i = tid; i < 256 * 64; i += 128) s.B[i] = 0.002f * (i % 89);
    if (tid == 0) { for (int i = 0; i < 64; ++i) mbar_init(&s.bar[i], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&s.tmem_base, 512);
    fence_proxy_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tb = s.tmem_base;
    if (tid == 0) {
        const uint32_t a = smem_u32(s.A), b = smem_u32(s.B);
        int cfg = 0;
        auto run = [&](auto issue) {
            issue(0); mma_commit(&s.bar[cfg]); mbar_wait(&s.bar[cfg], 0); fence_after_sync();
            long long t0 = clock64();
            for (int r = 0; r < R; ++r) issue(r);
            mma_commit(&s.bar[cfg]);
            mbar_wait(&s.bar[cfg], 1); fence_after_sync();
            long long t2 = clock64();
            out[cfg] = t2 - t0; ++cfg;
        };
        const int Ns[4] = {32, 48, 64, 128};
        const int Gs[3] = {1, 2, 4};
        for (int ts = 0; ts < 2; ++ts)
            for (int ni = 0; ni < 4; ++ni)
                for (int gi = 0; gi < 3; ++gi) {
                    const int N = Ns[ni], G = Gs[gi];
                    if (ts == 1 && N * G > 256) { out[cfg++] = 0; continue; }     // A operand lives in columns [0,128)
                    run([&](int r) {
                        const uint64_t bd = make_smem_desc(b + chunk_off(N, 0, 2 * (r % 6)), N * 16, 128);
                        if (ts == 0) mma_tf32_ss(tb + (r % G) * N, make_smem_desc(a + chunk_off(128, 0, 2 * (r % 6)), 2048, 128), bd, make_idesc_tf32(128, N), true);
                        else mma_tf32_ts(tb + 256 + (r % G) * N, tb + 8 * (r % 16), bd, make_idesc_tf32(128, N), true);
                    });
                }
    }
    fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 64 * 8); cudaMemset(d, 0, 64 * 8);
    cudaFuncSetAttribute(timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    const int R = 192;
    timing_kernel<<<1, 128, sizeof(Smem)>>>(d, R);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const int Ns[4] = {32, 48, 64, 128}; const int Gs[3] = {1, 2, 4};
    int c = 0;
    for (int ts = 0; ts < 2; ++ts) for (int ni = 0; ni < 4; ++ni) for (int gi = 0; gi < 3; ++gi, ++c)
        printf("tf32 %s M128 N%-3d K8, %d accumulators: %6.1f cyc/MMA\n", ts ? "TS" : "SS", Ns[ni], Gs[gi], (double)h[c] / R);
    return 0;
}
