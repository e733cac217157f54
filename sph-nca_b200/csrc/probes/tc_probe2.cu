// Probe 2: MN-major ("transposed") views of K-major chunk-layout tiles, as needed by the tensor-core
// backward.  Every tile is stored ONCE in the layout of tc_common.cuh (chunk (row m, c=k/4) at
// (c*R + m)*16 B) and consumed through two descriptor forms.  Single-pass tf32 (exactness of the
// addressing is what is probed); reference uses tf32-rounded inputs.
//   T4  D[128p x 48k]  = A_tmem[128p x 128h] * B,  B[n=k][K=h] = W1[h][k]   (W1 tile rows=h, MN-major B)
//   T5  D[128p x 128h] = A_smem[128p x 32o]  * B,  B[n=h][K=o] = W2[o][h]   (W2 tile rows=o, MN-major B)
//   T6  D[128h x 64p]  = A * B,  A[m=h][K=o] = W2[o][h] (MN-major A from smem),  B[n=p][K=o] = dO[p][o] (K-major)
//   T7  D[128h x 64k]  = A_tmem[128h x 64p] * B,  B[n=k][K=p] = y[p][k]     (y tile rows=p, MN-major B)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../tc_common.cuh"
using namespace sphnca::tc;

struct Smem {
    float W1[128 * 48];   // rows h (128), K=48
    float W2[32 * 128];   // rows o (32), K=128 (h)
    float dO[128 * 32];   // rows p (128), K=32 (o)
    float y[64 * 64];     // rows p (64), K=64 (k)
    uint64_t bar[4];
    uint32_t tmem_base;
};
__host__ __device__ constexpr uint32_t idesc_tf32_major(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
    return make_idesc_tf32(M, N) | (a_mn << 15) | (b_mn << 16);
}
__device__ void fill(float* t, const float* g, int rows, int K, int tid) {
    for (int e = tid; e < rows * K; e += 128) { int m = e / K, k = e % K; float h, l; split_tf32(g[e], h, l); t[chunk_off(rows, m, k / 4) / 4 + (k % 4)] = h; }
}
__global__ void __launch_bounds__(128) probe2(const float* W1, const float* W2, const float* dO, const float* y, const float* At4, const float* At7,
                                              float* D4, float* D5, float* D6, float* D7) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&s.bar[i], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&s.tmem_base, 512);
    fill(s.W1, W1, 128, 48, tid); fill(s.W2, W2, 32, 128, tid); fill(s.dO, dO, 128, 32, tid); fill(s.y, y, 64, 64, tid);
    fence_proxy_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tb = s.tmem_base, lane_addr = tb + ((uint32_t)(warp * 32) << 16);
    // stage TMEM A operands: T4 A[128p x 128h] at cols 0..127 ; T7 A[128h x 64p] at cols 128..191 (tf32-rounded)
    for (int c = 0; c < 128; c += 16) { uint32_t r[16]; for (int j = 0; j < 16; ++j) { float h, l; split_tf32(At4[tid * 128 + c + j], h, l); r[j] = __float_as_uint(h); } tmem_st16(lane_addr + c, r); }
    for (int c = 0; c < 64; c += 16) { uint32_t r[16]; for (int j = 0; j < 16; ++j) { float h, l; split_tf32(At7[tid * 64 + c + j], h, l); r[j] = __float_as_uint(h); } tmem_st16(lane_addr + 128 + c, r); }
    wait_st(); fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t cD4 = 192, cD5 = 256, cD6 = 384, cD7 = 448;
    if (tid == 0) {
        // T4: K = h (128) in 16 steps of 8; B = W1 tile MN-major: SBO (next 4 k) = 128*16, LBO (next 8 h) = 128
        for (int ks = 0; ks < 16; ++ks)
            mma_tf32_ts(tb + cD4, tb + 8 * ks, make_smem_desc(smem_u32(s.W1) + ks * 128, 128, 128 * 16), idesc_tf32_major(128, 48, 0, 1), ks > 0);
        // T5: K = o (32) in 4 steps; A = dO tile K-major (rows 128); B = W2 tile MN-major: SBO (next 4 h) = 32*16, LBO (next 8 o) = 128
        for (int ks = 0; ks < 4; ++ks)
            mma_tf32_ss(tb + cD5, make_smem_desc(smem_u32(s.dO) + chunk_off(128, 0, 2 * ks), 128 * 16, 128),
                        make_smem_desc(smem_u32(s.W2) + ks * 128, 128, 32 * 16), idesc_tf32_major(128, 128, 0, 1), ks > 0);
        // T6: A = W2 tile MN-major (M = h), B = dO tile rows 0..63 K-major (N = 64 p)
        for (int ks = 0; ks < 4; ++ks)
            mma_tf32_ss(tb + cD6, make_smem_desc(smem_u32(s.W2) + ks * 128, 128, 32 * 16),
                        make_smem_desc(smem_u32(s.dO) + chunk_off(128, 0, 2 * ks), 128 * 16, 128), idesc_tf32_major(128, 64, 1, 0), ks > 0);
        // T7: A from TMEM [128h x 64p]; B = y tile MN-major: SBO (next 4 k) = 64*16, LBO (next 8 p) = 128 ; K = p (64) in 8 steps
        for (int ks = 0; ks < 8; ++ks)
            mma_tf32_ts(tb + cD7, tb + 128 + 8 * ks, make_smem_desc(smem_u32(s.y) + ks * 128, 128, 64 * 16), idesc_tf32_major(128, 64, 0, 1), ks > 0);
        mma_commit(&s.bar[0]);
    }
    mbar_wait(&s.bar[0], 0); fence_after_sync();
    auto dump = [&](uint32_t col, int n, float* out) { for (int c = 0; c < n; c += 16) { uint32_t r[16]; tmem_ld16(lane_addr + col + c, r); wait_ld(); for (int j = 0; j < 16; ++j) out[(size_t)tid * n + c + j] = __uint_as_float(r[j]); } };
    dump(cD4, 48, D4); dump(cD5, 128, D5); dump(cD6, 64, D6); dump(cD7, 64, D7);
    fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}
static float tf(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }
int main() {
    srand(2); auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    std::vector<float> W1(128 * 48), W2(32 * 128), dO(128 * 32), y(64 * 64), A4(128 * 128), A7(128 * 64);
    for (auto* v : {&W1, &W2, &dO, &y, &A4, &A7}) for (auto& e : *v) e = rnd();
    float *d[6], *o[4]; std::vector<float>* src[6] = {&W1, &W2, &dO, &y, &A4, &A7}; int on[4] = {128 * 48, 128 * 128, 128 * 64, 128 * 64};
    for (int i = 0; i < 6; ++i) { cudaMalloc(&d[i], src[i]->size() * 4); cudaMemcpy(d[i], src[i]->data(), src[i]->size() * 4, cudaMemcpyHostToDevice); }
    for (int i = 0; i < 4; ++i) cudaMalloc(&o[i], on[i] * 4);
    cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    probe2<<<1, 128, sizeof(Smem)>>>(d[0], d[1], d[2], d[3], d[4], d[5], o[0], o[1], o[2], o[3]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<float> D4(on[0]), D5(on[1]), D6(on[2]), D7(on[3]);
    cudaMemcpy(D4.data(), o[0], on[0] * 4, cudaMemcpyDeviceToHost); cudaMemcpy(D5.data(), o[1], on[1] * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D6.data(), o[2], on[2] * 4, cudaMemcpyDeviceToHost); cudaMemcpy(D7.data(), o[3], on[3] * 4, cudaMemcpyDeviceToHost);
    double e4 = 0, e5 = 0, e6 = 0, e7 = 0, m4 = 0, m5 = 0, m6 = 0, m7 = 0;
    for (int p = 0; p < 128; ++p) for (int k = 0; k < 48; ++k) { double r = 0; for (int h = 0; h < 128; ++h) r += (double)tf(A4[p * 128 + h]) * tf(W1[h * 48 + k]); e4 = fmax(e4, fabs(D4[p * 48 + k] - r)); m4 = fmax(m4, fabs(r)); }
    for (int p = 0; p < 128; ++p) for (int h = 0; h < 128; ++h) { double r = 0; for (int q = 0; q < 32; ++q) r += (double)tf(dO[p * 32 + q]) * tf(W2[q * 128 + h]); e5 = fmax(e5, fabs(D5[p * 128 + h] - r)); m5 = fmax(m5, fabs(r)); }
    for (int h = 0; h < 128; ++h) for (int p = 0; p < 64; ++p) { double r = 0; for (int q = 0; q < 32; ++q) r += (double)tf(W2[q * 128 + h]) * tf(dO[p * 32 + q]); e6 = fmax(e6, fabs(D6[h * 64 + p] - r)); m6 = fmax(m6, fabs(r)); }
    for (int h = 0; h < 128; ++h) for (int k = 0; k < 64; ++k) { double r = 0; for (int p = 0; p < 64; ++p) r += (double)tf(A7[h * 64 + p]) * tf(y[p * 64 + k]); e7 = fmax(e7, fabs(D7[h * 64 + k] - r)); m7 = fmax(m7, fabs(r)); }
    printf("T4 TS, B MN-major (W1 view)      rel err %.3e\nT5 SS, B MN-major (W2 view)      rel err %.3e\nT6 SS, A MN-major (W2 view)      rel err %.3e\nT7 TS transposed, B MN-major (y) rel err %.3e\n", e4 / m4, e5 / m5, e6 / m6, e7 / m7);
    bool ok = e4 / m4 < 1e-5 && e5 / m5 < 1e-5 && e6 / m6 < 1e-5 && e7 / m7 < 1e-5;
    printf(ok ? "PROBE2 OK\n" : "PROBE2 FAILED\n"); return ok ? 0 : 1;
}
