// Stand-alone probe for the tcgen05 building blocks of the fused MLP (run on a B200 via gpurun):
//   T1  D1[128x128] = A[128x48] * B1[128x48]^T   kind::tf32, operands in smem (K-major, no swizzle)
//   T2  same with the 3xTF32 split (hi*hi + hi*lo + lo*hi) -> fp32-class accuracy vs fp64
//   T3  D2[128x32]  = H[128x128] * B2[32x128]^T  with A = H staged in TMEM by tcgen05.st (3xTF32)
// Prints max relative errors; exits non-zero on failure.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../tc_common.cuh"

using namespace sphnca::tc;

constexpr int M = 128, K1 = 48, N1 = 128, K2 = 128, N2 = 32;

struct Smem {
    float A_hi[M * K1], A_lo[M * K1];
    float B1_hi[N1 * K1], B1_lo[N1 * K1];
    float B2_hi[N2 * K2], B2_lo[N2 * K2];
    uint64_t bar[2];
    uint32_t tmem_base;
};

__device__ void fill_tile(float* hi, float* lo, const float* g, int rows, int K, int tid, int nthreads) {
    for (int e = tid; e < rows * K; e += nthreads) {
        int m = e / K, k = e % K;
        float h, l;
        split_tf32(g[e], h, l);
        uint32_t off = chunk_off(rows, m, k / 4) / 4 + (k % 4);
        hi[off] = h; lo[off] = l;
    }
}

__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* B1, const float* H, const float* B2,
                                                    float* D_tf32, float* D_3x, float* D2) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&s.bar[0], 1); mbar_init(&s.bar[1], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&s.tmem_base, 512);
    fill_tile(s.A_hi, s.A_lo, A, M, K1, tid, 128);
    fill_tile(s.B1_hi, s.B1_lo, B1, N1, K1, tid, 128);
    fill_tile(s.B2_hi, s.B2_lo, B2, N2, K2, tid, 128);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = s.tmem_base;
    const uint32_t col_D1 = 0, col_D3x = 128, col_H = 256 /* hi 256..383 */, col_D2 = 448;   // H_lo reuses col 0..127 later
    const uint32_t idesc1 = make_idesc_tf32(M, N1), idesc2 = make_idesc_tf32(M, N2);
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);

    // ---- T1 + T2 ----
    if (tid == 0) {
        for (int ks = 0; ks < K1 / 8; ++ks) {
            uint64_t a_hi = make_smem_desc(smem_u32(s.A_hi) + chunk_off(M, 0, 2 * ks), M * 16, 128);
            uint64_t a_lo = make_smem_desc(smem_u32(s.A_lo) + chunk_off(M, 0, 2 * ks), M * 16, 128);
            uint64_t b_hi = make_smem_desc(smem_u32(s.B1_hi) + chunk_off(N1, 0, 2 * ks), N1 * 16, 128);
            uint64_t b_lo = make_smem_desc(smem_u32(s.B1_lo) + chunk_off(N1, 0, 2 * ks), N1 * 16, 128);
            mma_tf32_ss(tbase + col_D1, a_hi, b_hi, idesc1, ks > 0);
            mma_tf32_ss(tbase + col_D3x, a_hi, b_hi, idesc1, ks > 0);
            mma_tf32_ss(tbase + col_D3x, a_hi, b_lo, idesc1, true);
            mma_tf32_ss(tbase + col_D3x, a_lo, b_hi, idesc1, true);
        }
        mma_commit(&s.bar[0]);
    }
    mbar_wait(&s.bar[0], 0);
    fence_after_sync();
    for (int c = 0; c < N1; c += 16) {
        uint32_t r[16];
        tmem_ld16(lane_addr + col_D1 + c, r); wait_ld();
        for (int j = 0; j < 16; ++j) D_tf32[(size_t)tid * N1 + c + j] = __uint_as_float(r[j]);
        tmem_ld16(lane_addr + col_D3x + c, r); wait_ld();
        for (int j = 0; j < 16; ++j) D_3x[(size_t)tid * N1 + c + j] = __uint_as_float(r[j]);
    }
    // ---- T3: stage H (row = thread) into TMEM as hi (cols 256..383) and lo (cols 0..127), then TS MMAs ----
    fence_before_sync();
    __syncthreads();          // everyone finished reading D1 (cols 0..127) before it is overwritten by H_lo
    fence_after_sync();
    const uint32_t col_Hlo = 0;
    for (int c = 0; c < K2; c += 16) {
        uint32_t hi[16], lo[16];
        for (int j = 0; j < 16; ++j) {
            float h, l; split_tf32(H[(size_t)tid * K2 + c + j], h, l);
            hi[j] = __float_as_uint(h); lo[j] = __float_as_uint(l);
        }
        tmem_st16(lane_addr + col_H + c, hi);
        tmem_st16(lane_addr + col_Hlo + c, lo);
    }
    wait_st();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid == 0) {
        for (int ks = 0; ks < K2 / 8; ++ks) {
            uint64_t b_hi = make_smem_desc(smem_u32(s.B2_hi) + chunk_off(N2, 0, 2 * ks), N2 * 16, 128);
            uint64_t b_lo = make_smem_desc(smem_u32(s.B2_lo) + chunk_off(N2, 0, 2 * ks), N2 * 16, 128);
            mma_tf32_ts(tbase + col_D2, tbase + col_H + 8 * ks, b_hi, idesc2, ks > 0);
            mma_tf32_ts(tbase + col_D2, tbase + col_H + 8 * ks, b_lo, idesc2, true);
            mma_tf32_ts(tbase + col_D2, tbase + col_Hlo + 8 * ks, b_hi, idesc2, true);
        }
        mma_commit(&s.bar[1]);
    }
    mbar_wait(&s.bar[1], 0);
    fence_after_sync();
    for (int c = 0; c < N2; c += 16) {
        uint32_t r[16];
        tmem_ld16(lane_addr + col_D2 + c, r); wait_ld();
        for (int j = 0; j < 16; ++j) D2[(size_t)tid * N2 + c + j] = __uint_as_float(r[j]);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

static float tf32_round(float x) {            // round-to-nearest (ties away), like cvt.rna.tf32.f32
    uint32_t u; memcpy(&u, &x, 4);
    u += 0x1000u; u &= 0xFFFFE000u;
    float r; memcpy(&r, &u, 4); return r;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

int main() {
    std::vector<float> A(M * K1), B1(N1 * K1), H(M * K2), B2(N2 * K2);
    srand(1);
    auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (auto& v : A) v = rnd(); for (auto& v : B1) v = rnd(); for (auto& v : H) v = fabsf(rnd()); for (auto& v : B2) v = rnd();
    float *dA, *dB1, *dH, *dB2, *dD, *dD3, *dD2;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB1, B1.size() * 4)); CK(cudaMalloc(&dH, H.size() * 4)); CK(cudaMalloc(&dB2, B2.size() * 4));
    CK(cudaMalloc(&dD, M * N1 * 4)); CK(cudaMalloc(&dD3, M * N1 * 4)); CK(cudaMalloc(&dD2, M * N2 * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB1, B1.data(), B1.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dH, H.data(), H.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB2, B2.data(), B2.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    probe_kernel<<<1, 128, sizeof(Smem)>>>(dA, dB1, dH, dB2, dD, dD3, dD2);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    std::vector<float> D(M * N1), D3(M * N1), D2(M * N2);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(D3.data(), dD3, D3.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost));
    double e1 = 0, e2 = 0, e3 = 0, mx1 = 0, mx3 = 0, e1x = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N1; ++n) {
        double rt = 0, rx = 0;
        for (int k = 0; k < K1; ++k) { rt += (double)tf32_round(A[m * K1 + k]) * tf32_round(B1[n * K1 + k]); rx += (double)A[m * K1 + k] * B1[n * K1 + k]; }
        e1 = fmax(e1, fabs(D[m * N1 + n] - rt)); e1x = fmax(e1x, fabs(D[m * N1 + n] - rx)); e2 = fmax(e2, fabs(D3[m * N1 + n] - rx)); mx1 = fmax(mx1, fabs(rx));
    }
    for (int m = 0; m < M; ++m) for (int n = 0; n < N2; ++n) {
        double rx = 0;
        for (int k = 0; k < K2; ++k) rx += (double)H[m * K2 + k] * B2[n * K2 + k];
        e3 = fmax(e3, fabs(D2[m * N2 + n] - rx)); mx3 = fmax(mx3, fabs(rx));
    }
    printf("T1 tf32 SS      : max|D - tf32ref| / max|ref| = %.3e   (vs exact: %.3e)\n", e1 / mx1, e1x / mx1);
    printf("T2 3xTF32 SS    : max|D - exact|   / max|ref| = %.3e\n", e2 / mx1);
    printf("T3 3xTF32 TS    : max|D - exact|   / max|ref| = %.3e\n", e3 / mx3);
    bool ok = e1 / mx1 < 1e-5 && e2 / mx1 < 2e-6 && e3 / mx3 < 2e-6;
    printf(ok ? "PROBE OK\n" : "PROBE FAILED\n");
    return ok ? 0 : 1;
}
