// Probe for round 2 (run on a B200): is a 3-pass FP16 split a drop-in for the 3xTF32 split of the MLP kernels?
//   x = hi + lo,  hi = fp16(x),  lo = fp16((x - hi) * 2^11)  (the scale keeps lo a NORMAL fp16 number whenever hi is one)
//   D = A_hi.B_hi + 2^-11 (A_hi.B_lo + A_lo.B_hi), the two brackets in separate accumulator columns, fp32 accumulation
// kind::f16 covers K = 16 per instruction (kind::tf32: K = 8) at the same operand bytes per MMA, so -- if the per-MMA cost at
// small N is the shared-memory operand fetch, as tc_timing4 says -- half the MMAs and half the tensor time.
// Prints the error of: single-pass tf32, 3xTF32, single-pass fp16, 3x fp16 (scaled lo), 3x fp16 (unscaled lo, one accumulator)
// against fp64, for operands of magnitude ~1 and ~1e-3, and the cycles per kind::f16 MMA for N = 32..256.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "../tc_common.cuh"

using namespace sphnca::tc;

constexpr int M = 128, N = 128, K = 64;
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t M_, uint32_t N_) { return (1u << 4) | ((N_ >> 3) << 17) | ((M_ >> 4) << 24); }

struct Smem {
    float At_hi[M * K], At_lo[M * K], Bt_hi[N * K], Bt_lo[N * K];          // tf32 containers
    __half Ah[M * K], Al[M * K], Alu[M * K], Bh[N * K], Bl[N * K], Blu[N * K];   // fp16: hi, scaled lo, unscaled lo
    uint64_t bar[16];
    uint32_t tmem_base;
};

__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, bool acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)acc) : "memory");
}

__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* out /* 6 x [M][N] */, long long* cyc) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { for (int i = 0; i < 16; ++i) mbar_init(&s.bar[i], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&s.tmem_base, 512);
    for (int e = tid; e < M * K; e += 128) {
        const int m = e / K, k = e % K;
        float h, l; split_tf32(A[e], h, l);
        const uint32_t o32 = chunk_off(M, m, k / 4) / 4 + (k % 4), o16 = chunk_off(M, m, k / 8) / 2 + (k % 8);
        s.At_hi[o32] = h; s.At_lo[o32] = l;
        const __half hh = __float2half_rn(A[e]); const float r = A[e] - __half2float(hh);
        s.Ah[o16] = hh; s.Al[o16] = __float2half_rn(r * 2048.f); s.Alu[o16] = __float2half_rn(r);
    }
    for (int e = tid; e < N * K; e += 128) {
        const int m = e / K, k = e % K;
        float h, l; split_tf32(B[e], h, l);
        const uint32_t o32 = chunk_off(N, m, k / 4) / 4 + (k % 4), o16 = chunk_off(N, m, k / 8) / 2 + (k % 8);
        s.Bt_hi[o32] = h; s.Bt_lo[o32] = l;
        const __half hh = __float2half_rn(B[e]); const float r = B[e] - __half2float(hh);
        s.Bh[o16] = hh; s.Bl[o16] = __float2half_rn(r * 2048.f); s.Blu[o16] = __float2half_rn(r);
    }
    fence_proxy_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tb = s.tmem_base;
    const uint32_t i32 = make_idesc_tf32(M, N), i16 = make_idesc_f16(M, N);
    auto d32 = [&](const float* p, int rows, int ks) { return make_smem_desc(smem_u32(p) + chunk_off(rows, 0, 2 * ks), rows * 16, 128); };
    auto d16 = [&](const __half* p, int rows, int ks) { return make_smem_desc(smem_u32(p) + chunk_off(rows, 0, 2 * ks), rows * 16, 128); };
    if (tid == 0) {
        // columns: 0 tf32 single | 128 3xTF32 | 256 fp16 single (= hi.hi) | 384 fp16 cross terms (scaled lo)
        for (int ks = 0; ks < K / 8; ++ks) {
            mma_tf32_ss(tb + 0, d32(s.At_hi, M, ks), d32(s.Bt_hi, N, ks), i32, ks > 0);
            mma_tf32_ss(tb + 128, d32(s.At_lo, M, ks), d32(s.Bt_hi, N, ks), i32, ks > 0);
            mma_tf32_ss(tb + 128, d32(s.At_hi, M, ks), d32(s.Bt_lo, N, ks), i32, true);
            mma_tf32_ss(tb + 128, d32(s.At_hi, M, ks), d32(s.Bt_hi, N, ks), i32, true);
        }
        for (int ks = 0; ks < K / 16; ++ks) {
            mma_f16_ss(tb + 256, d16(s.Ah, M, ks), d16(s.Bh, N, ks), i16, ks > 0);
            mma_f16_ss(tb + 384, d16(s.Ah, M, ks), d16(s.Bl, N, ks), i16, ks > 0);
            mma_f16_ss(tb + 384, d16(s.Al, M, ks), d16(s.Bh, N, ks), i16, true);
        }
        mma_commit(&s.bar[0]);
    }
    mbar_wait(&s.bar[0], 0); fence_after_sync();
    const uint32_t lane_addr = tb + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < N; c += 16) {
        uint32_t r0[16], r1[16], r2[16], r3[16];
        tmem_ld16(lane_addr + c, r0); tmem_ld16(lane_addr + 128 + c, r1); tmem_ld16(lane_addr + 256 + c, r2); tmem_ld16(lane_addr + 384 + c, r3); wait_ld();
        for (int j = 0; j < 16; ++j) {
            const size_t o = (size_t)tid * N + c + j;
            out[o] = __uint_as_float(r0[j]); out[M * N + o] = __uint_as_float(r1[j]); out[2 * M * N + o] = __uint_as_float(r2[j]);
            out[3 * M * N + o] = fmaf(__uint_as_float(r3[j]), 1.f / 2048.f, __uint_as_float(r2[j]));
        }
    }
    fence_before_sync(); __syncthreads(); fence_after_sync();
    if (tid == 0) {       // unscaled lo, all three passes into one accumulator (columns 0..127 again)
        for (int ks = 0; ks < K / 16; ++ks) {
            mma_f16_ss(tb + 0, d16(s.Alu, M, ks), d16(s.Bh, N, ks), i16, ks > 0);
            mma_f16_ss(tb + 0, d16(s.Ah, M, ks), d16(s.Blu, N, ks), i16, true);
            mma_f16_ss(tb + 0, d16(s.Ah, M, ks), d16(s.Bh, N, ks), i16, true);
        }
        mma_commit(&s.bar[1]);
    }
    mbar_wait(&s.bar[1], 0); fence_after_sync();
    for (int c = 0; c < N; c += 16) {
        uint32_t r0[16];
        tmem_ld16(lane_addr + c, r0); wait_ld();
        for (int j = 0; j < 16; ++j) out[4 * M * N + (size_t)tid * N + c + j] = __uint_as_float(r0[j]);
    }
    fence_before_sync(); __syncthreads(); fence_after_sync();
    // ---- timing: 96 kind::f16 MMAs (M128, K16) per N, issue predicated on the elected lane (uniform operands)
    if (warp == 0) {
        const uint32_t go = elect_one();
        int cfg = 0;
        auto run = [&](const int n_) {
            const uint32_t id = make_idesc_f16(M, n_);
            auto issue = [&]() {
#pragma unroll 8
                for (int r = 0; r < 96; ++r) {
                    const uint64_t a = d16(s.Ah, M, r % 4), b = make_smem_desc(smem_u32(s.Bh) + chunk_off(n_, 0, 2 * (r % 4)), n_ * 16, 128);
                    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;\n\t}\n"
                                 :: "r"(tb + 256), "l"(a), "l"(b), "r"(id), "r"(go) : "memory");
                }
            };
            issue(); if (go) mma_commit(&s.bar[2 + cfg]); __syncwarp(); mbar_wait(&s.bar[2 + cfg], 0); fence_after_sync();
            const long long t0 = clock64();
            issue();
            if (go) mma_commit(&s.bar[2 + cfg]);
            __syncwarp();
            mbar_wait(&s.bar[2 + cfg], 1); fence_after_sync();
            const long long t2 = clock64();
            if (go) cyc[cfg] = t2 - t0;
            ++cfg;
        };
        run(32); run(48); run(64); run(128); run(256);
    }
    fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

int main() {
    float *dA, *dB, *dO; long long* dC;
    CK(cudaMalloc(&dA, M * K * 4)); CK(cudaMalloc(&dB, N * K * 4)); CK(cudaMalloc(&dO, 6 * M * N * 4)); CK(cudaMalloc(&dC, 64));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    const char* nm[5] = {"tf32 single pass", "3xTF32", "fp16 single pass", "3xFP16, lo scaled by 2^11", "3xFP16, lo unscaled"};
    for (int cs = 0; cs < 3; ++cs) {
        const float magA = cs == 0 ? 1.f : (cs == 1 ? 1e-3f : 30.f), magB = cs == 2 ? 1e-3f : magA;
        std::vector<float> A(M * K), B(N * K);
        srand(7 + cs);
        for (auto& v : A) v = magA * ((float)rand() / RAND_MAX * 2.f - 1.f);
        for (auto& v : B) v = magB * ((float)rand() / RAND_MAX * 2.f - 1.f);
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(dC, 0, 64));
        probe<<<1, 128, sizeof(Smem)>>>(dA, dB, dO, dC);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        std::vector<float> O(6 * M * N); long long cyc[8];
        CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(cyc, dC, 64, cudaMemcpyDeviceToHost));
        double mx = 0; std::vector<double> ref(M * N);
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double acc = 0; for (int k = 0; k < K; ++k) acc += (double)A[m * K + k] * (double)B[n * K + k];
            ref[m * N + n] = acc; mx = fmax(mx, fabs(acc));
        }
        printf("operands ~%g x ~%g (K = %d): max |D - fp64| / max |D|\n", magA, magB, K);
        for (int v = 0; v < 5; ++v) {
            double e = 0; for (int i = 0; i < M * N; ++i) e = fmax(e, fabs((double)O[(size_t)v * M * N + i] - ref[i]));
            printf("   %-28s %.3e\n", nm[v], e / mx);
        }
        if (cs == 0) {
            const int ns[5] = {32, 48, 64, 128, 256};
            for (int c = 0; c < 5; ++c) printf("f16 SS N%-4d M128 K16 x96: %6.1f cyc/MMA (kind::tf32 K8: see tc_timing4)\n", ns[c], cyc[c] / 96.0);
        }
    }
    return 0;
}
