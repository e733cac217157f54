// Probe 3: find the descriptor form for an MN-major tf32 B operand (no swizzle).
//   D[128p x 128h] = A_smem[128p x 32o] * B,  B[n=h][K=o] = W2[o][h], W2 tile stored rows=o (32), K=h(128) chunk layout.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../tc_common.cuh"
using namespace sphnca::tc;
struct Smem { float W2[32 * 128]; float W2b[128 * 32]; float dO[128 * 32]; uint64_t bar[16]; uint32_t tmem_base; };
__device__ void fill(float* t, const float* g, int rows, int K, int tid) {
    for (int e = tid; e < rows * K; e += 128) { int m = e / K, k = e % K; float h, l; split_tf32(g[e], h, l); t[chunk_off(rows, m, k / 4) / 4 + (k % 4)] = h; }
}
__device__ uint64_t desc_raw(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lbo_mode) {
    return make_smem_desc(saddr, lbo, sbo) | ((uint64_t)lbo_mode << 52);
}
__global__ void __launch_bounds__(128) probe3(const float* W2, const float* dO, float* out, int variant) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&s.bar[0], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&s.tmem_base, 512);
    fill(s.W2, W2, 32, 128, tid); fill(s.dO, dO, 128, 32, tid);
    // W2b: an explicitly MN-major tile: element (h, o) at ((o/8)*32 + h/4)*128 + (o%8)*16 + (h%4)*4  [core matrix = 8 o-rows x 4 h]
    for (int e = tid; e < 32 * 128; e += 128) { int o = e / 128, h = e % 128; float hi, lo; split_tf32(W2[e], hi, lo); s.W2b[(((o / 8) * 32 + h / 4) * 128 + (o % 8) * 16 + (h % 4) * 4) / 4] = hi; }
    fence_proxy_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tb = s.tmem_base, lane_addr = tb + ((uint32_t)(warp * 32) << 16);
    const uint32_t idesc = make_idesc_tf32(128, 128) | (1u << 16);
    if (tid == 0) {
        for (int ks = 0; ks < 4; ++ks) {
            uint64_t a = make_smem_desc(smem_u32(s.dO) + chunk_off(128, 0, 2 * ks), 128 * 16, 128);
            uint64_t b;
            switch (variant) {
                case 0: b = desc_raw(smem_u32(s.W2) + ks * 128, 128, 512, 0); break;          // as designed: LBO=next 8 o, SBO=next 4 h
                case 1: b = desc_raw(smem_u32(s.W2) + ks * 128, 512, 128, 0); break;          // swapped
                case 2: b = desc_raw(smem_u32(s.W2b) + ks * 4096, 4096, 128, 0); break;       // explicit tile: next-4h = 128 B, next-8o = 4096 B ; LBO=4096,SBO=128
                case 3: b = desc_raw(smem_u32(s.W2b) + ks * 4096, 128, 4096, 0); break;       // swapped
                case 4: b = desc_raw(smem_u32(s.W2) + ks * 128, 128, 512, 1); break;
                case 5: b = desc_raw(smem_u32(s.W2) + ks * 128, 512, 128, 1); break;
                case 6: b = desc_raw(smem_u32(s.W2b) + ks * 4096, 4096, 128, 1); break;
                default: b = desc_raw(smem_u32(s.W2b) + ks * 4096, 128, 4096, 1); break;
            }
            mma_tf32_ss(tb, a, b, idesc, ks > 0);
        }
        mma_commit(&s.bar[0]);
    }
    mbar_wait(&s.bar[0], 0); fence_after_sync();
    for (int c = 0; c < 128; c += 16) { uint32_t r[16]; tmem_ld16(lane_addr + c, r); wait_ld(); for (int j = 0; j < 16; ++j) out[(size_t)tid * 128 + c + j] = __uint_as_float(r[j]); }
    fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}
static float tf(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }
int main() {
    srand(2); auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    std::vector<float> W2(32 * 128), dO(128 * 32); for (auto& e : W2) e = rnd(); for (auto& e : dO) e = rnd();
    float *dW, *dD, *o; cudaMalloc(&dW, W2.size() * 4); cudaMalloc(&dD, dO.size() * 4); cudaMalloc(&o, 128 * 128 * 4);
    cudaMemcpy(dW, W2.data(), W2.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dD, dO.data(), dO.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    std::vector<double> ref(128 * 128); double mx = 0;
    for (int p = 0; p < 128; ++p) for (int h = 0; h < 128; ++h) { double r = 0; for (int q = 0; q < 32; ++q) r += (double)tf(dO[p * 32 + q]) * tf(W2[q * 128 + h]); ref[p * 128 + h] = r; mx = fmax(mx, fabs(r)); }
    for (int v = 0; v < 8; ++v) {
        cudaMemset(o, 0, 128 * 128 * 4);
        probe3<<<1, 128, sizeof(Smem)>>>(dW, dD, o, v);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: CUDA error: %s\n", v, cudaGetErrorString(e)); return 2; }
        std::vector<float> D(128 * 128); cudaMemcpy(D.data(), o, D.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0; int nz = 0; for (int i = 0; i < 128 * 128; ++i) { err = fmax(err, fabs(D[i] - ref[i])); nz += D[i] != 0.f; }
        printf("variant %d: rel err %.3e  nonzero %d  D[0..3] = %g %g %g %g   ref %g %g %g %g\n", v, err / mx, nz, D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3]);
    }
    return 0;
}
