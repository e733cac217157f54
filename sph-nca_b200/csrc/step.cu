// Fused NCA step (reference code/nca.py:87-117), 16 channels, hidden 256.
//
// Forward = 2 kernels (the reference: 3 Numba sweeps + ~25 eager torch kernels + 2 cuBLAS GEMMs):
//   step_fwd_kernel   one sweep over the 3^D cells computes BOTH the previous life mask
//                     (blur of alive bits, F=1) and the SPH gradient (x,y components, F=16);
//                     the 48-wide perception vector stays in registers and feeds the
//                     48->256->n_out MLP (weights staged once per CTA in shared memory,
//                     broadcast LDS.128), then the gated/orig update + fire mask.
//   step_mask_kernel  second life-mask sweep over the NEW alive bits (needs all neighbours'
//                     updated state -> grid-wide dependency -> separate launch), applies
//                     prev & new to the state.
// Backward = 2 kernels: step_bwd_kernel (gate/tanh derivatives, MLP data gradient with recomputed
// hidden layer) and step_adjoint_kernel (gather adjoint of the gradient sweep; deterministic).
//
// All arrays are in cell-sorted order.  v1: the MLP runs on the fp32 CUDA-core pipe.
#include "step_common.cuh"
#include "sweep_tile.cuh"
#include "sweep_stage.cuh"

namespace sphnca {

struct MlpSmem {
    float W1[NH][NIN];       // row n = weights of hidden unit n (48 KB)
    float W2t[NH][NO_PAD];   // row n = outgoing weights of hidden unit n (36 KB)
    float b1[NH];
    float b2[NO_PAD];
};

__device__ __forceinline__ void stage_weights(MlpSmem& s, const sphnca_step_params& p, int n_out) {
    const float4* w1 = reinterpret_cast<const float4*>(p.W1);
    float4* d1 = reinterpret_cast<float4*>(&s.W1[0][0]);
    for (int i = threadIdx.x; i < NH * NIN / 4; i += blockDim.x) d1[i] = __ldg(w1 + i);
    for (int i = threadIdx.x; i < NH * NO_PAD; i += blockDim.x) {
        const int n = i / NO_PAD, o = i - n * NO_PAD;
        s.W2t[n][o] = o < n_out ? __ldg(p.W2 + (size_t)o * NH + n) : 0.f;
    }
    for (int i = threadIdx.x; i < NH; i += blockDim.x) s.b1[i] = __ldg(p.b1 + i);
    if (threadIdx.x < NO_PAD) s.b2[threadIdx.x] = threadIdx.x < n_out ? __ldg(p.b2 + threadIdx.x) : 0.f;
}

// hidden pre-activation of unit n for perception vector y (48 FMA, broadcast smem reads)
__device__ __forceinline__ float hidden_pre(const MlpSmem& s, int n, const float (&y)[NIN]) {
    const float4* w = reinterpret_cast<const float4*>(&s.W1[n][0]);
    float a0 = s.b1[n], a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int k = 0; k < NIN / 4; ++k) {
        const float4 t = w[k];
        a0 = fmaf(t.x, y[4*k], a0); a1 = fmaf(t.y, y[4*k+1], a1);
        a2 = fmaf(t.z, y[4*k+2], a2); a3 = fmaf(t.w, y[4*k+3], a3);
    }
    return (a0 + a1) + (a2 + a3);
}

template <int D>
__global__ void __launch_bounds__(kStepThreads, 2)
step_fwd_kernel(GridView g, const float4* __restrict__ xv4, const float* __restrict__ A_s,
                const uint8_t* __restrict__ fire_s, sphnca_step_params prm, StepConsts c,
                float* __restrict__ An_s, float* __restrict__ Y_s, float* __restrict__ O_s,
                uint8_t* __restrict__ flags_s, uint8_t* __restrict__ alive_s, uint32_t* __restrict__ relu_s) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MlpSmem& s = *reinterpret_cast<MlpSmem*>(smem_raw);
    stage_weights(s, prm, c.n_out);
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.n) return;

    // ---- perception sweep: gradient (x,y) of 16 channels + blur of the alive bit ----
    const float4 xi = xv4[p];
    float y[NIN];
    {
        float Ai[NF];
        load16(A_s + (size_t)p * NF, Ai);
#pragma unroll
        for (int f = 0; f < NF; ++f) { y[f] = Ai[f]; y[NF + f] = 0.f; y[2 * NF + f] = 0.f; }
    }
    float blur = 0.f;
    for_each_candidate<D>(g, g.sorted_cell[p], [&](int q) {
        float rx, ry;
        const float4 xj = __ldg(&xv4[q]);
        const float d2 = pair_d2<D>(xi, xj, rx, ry);
        if (d2 < c.h2) {
            float Aj[NF];
            load16(A_s + (size_t)q * NF, Aj);
            const float t = c.h2 - d2;
            const float alive = c.use_alpha ? (Aj[3] > 0.1f ? 1.f : 0.f) : 1.f;
            blur = fmaf(alive * xj.w, t * t * t, blur);
            if (d2 > 0.f) {
                const float d = sqrtf(d2);
                const float hd = c.h - d;
                const float m = 3.f * hd * hd / d * xj.w;
                const float wx = m * rx, wy = m * ry;
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    const float dA = Aj[f] - y[f];
                    y[NF + f] = fmaf(dA, wx, y[NF + f]);
                    y[2 * NF + f] = fmaf(dA, wy, y[2 * NF + f]);
                }
            }
        }
    });
    const bool prev_mask = blur * c.cw > 0.1f;
#pragma unroll
    for (int f = 0; f < 2 * NF; ++f) y[NF + f] *= c.kg;

    // ---- MLP 48 -> 256 (ReLU) -> n_out, one particle per thread ----
    float out[NO_PAD];
#pragma unroll
    for (int o = 0; o < NO_PAD; ++o) out[o] = s.b2[o];
    uint32_t* rm = relu_s + (size_t)p * (NH / 32);
#pragma unroll 1
    for (int nw = 0; nw < NH / 32; ++nw) {
        uint32_t bits = 0;
#pragma unroll 2
        for (int nb = 0; nb < 32; ++nb) {
            const int n = nw * 32 + nb;
            const float pre = hidden_pre(s, n, y);
            bits |= (pre > 0.f ? 1u : 0u) << nb;
            const float hval = fmaxf(pre, 0.f);
            const float4* w2 = reinterpret_cast<const float4*>(&s.W2t[n][0]);
#pragma unroll
            for (int k = 0; k < NO_PAD / 4; ++k) {
                const float4 t = w2[k];
                out[4*k] = fmaf(t.x, hval, out[4*k]); out[4*k+1] = fmaf(t.y, hval, out[4*k+1]);
                out[4*k+2] = fmaf(t.z, hval, out[4*k+2]); out[4*k+3] = fmaf(t.w, hval, out[4*k+3]);
            }
        }
        rm[nw] = bits;        // ReLU activity bits of hidden units 32*nw .. 32*nw+31 (saved for backward)
    }

    // ---- update rule + stochastic fire mask (nca.py:100-110) ----
    const bool fire = fire_s[p] != 0;
    float nA[NF];
    if (c.rule == 0) {
        const float mult = sigmoidf_(out[2 * NF]);
#pragma unroll
        for (int f = 0; f < NF; ++f) nA[f] = fmaf(y[f], sigmoidf_(out[f]), tanhf(out[NF + f]) * mult);
    } else {
#pragma unroll
        for (int f = 0; f < NF; ++f) nA[f] = fmaf(out[f], c.orig_rate, y[f]);
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) nA[f] = fire ? nA[f] : y[f];

    float4* o4 = reinterpret_cast<float4*>(An_s + (size_t)p * NF);
#pragma unroll
    for (int k = 0; k < 4; ++k) o4[k] = make_float4(nA[4*k], nA[4*k+1], nA[4*k+2], nA[4*k+3]);
    float4* y4 = reinterpret_cast<float4*>(Y_s + (size_t)p * 2 * NF);
#pragma unroll
    for (int k = 0; k < 8; ++k) y4[k] = make_float4(y[NF+4*k], y[NF+4*k+1], y[NF+4*k+2], y[NF+4*k+3]);
    float4* oo4 = reinterpret_cast<float4*>(O_s + (size_t)p * NO_PAD);
#pragma unroll
    for (int k = 0; k < NO_PAD / 4; ++k) oo4[k] = make_float4(out[4*k], out[4*k+1], out[4*k+2], out[4*k+3]);
    flags_s[p] = (uint8_t)((prev_mask ? 1 : 0) | (fire ? 4 : 0));
    alive_s[p] = c.use_alpha ? (nA[3] > 0.1f ? 1 : 0) : 1;
}

// Perception sweep only (split forward path): SPH gradient (x,y; 16 channels, scaled) + previous life mask.
// No shared memory -> the whole 228 KB of the SM serve as L1 for the neighbour gathers; 128-thread CTAs.
template <int D>
__global__ void __launch_bounds__(128)
step_perceive_kernel(GridView g, const float4* __restrict__ xv4, const float* __restrict__ A_s, StepConsts c,
                     float* __restrict__ Y_s, uint8_t* __restrict__ flags_s) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.n) return;
    const float4 xi = xv4[p];
    float2 Ai[NF / 2], gx[NF / 2], gy[NF / 2];            // channel pairs (FFMA2 / FADD2)
    {
        const float4* a4 = reinterpret_cast<const float4*>(A_s + (size_t)p * NF);
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float4 t = __ldg(a4 + k); Ai[2*k] = make_float2(t.x, t.y); Ai[2*k+1] = make_float2(t.z, t.w); }
    }
#pragma unroll
    for (int f = 0; f < NF / 2; ++f) { gx[f] = make_float2(0.f, 0.f); gy[f] = make_float2(0.f, 0.f); }
    float blur = 0.f;
    for_each_candidate<D>(g, g.sorted_cell[p], [&](int q) {
        float rx, ry;
        const float4 xj = __ldg(&xv4[q]);
        const float d2 = pair_d2<D>(xi, xj, rx, ry);
        if (d2 < c.h2) {
            const float4* a4 = reinterpret_cast<const float4*>(A_s + (size_t)q * NF);
            float2 Aj[NF / 2];
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float4 t = __ldg(a4 + k); Aj[2*k] = make_float2(t.x, t.y); Aj[2*k+1] = make_float2(t.z, t.w); }
            const float t = c.h2 - d2;
            const float alive = c.use_alpha ? (Aj[1].y > 0.1f ? 1.f : 0.f) : 1.f;      // channel 3
            blur = fmaf(alive * xj.w, t * t * t, blur);
            if (d2 > 0.f) {
                // 3 (h-d)^2 / d with d = d2 * rsqrt(d2): MUFU.RSQ (2 ulp) instead of IEEE sqrt + divide
                const float inv = rsqrtf(d2);
                const float hd = c.h - d2 * inv;
                const float m = 3.f * hd * hd * inv * xj.w;
                const float2 wx = make_float2(m * rx, m * rx), wy = make_float2(m * ry, m * ry);
#pragma unroll
                for (int f = 0; f < NF / 2; ++f) {
                    const float2 dA = fsub2(Aj[f], Ai[f]);
                    gx[f] = ffma2(dA, wx, gx[f]);
                    gy[f] = ffma2(dA, wy, gy[f]);
                }
            }
        }
    });
    float4* y4 = reinterpret_cast<float4*>(Y_s + (size_t)p * 2 * NF);
#pragma unroll
    for (int k = 0; k < 4; ++k) y4[k] = make_float4(gx[2*k].x * c.kg, gx[2*k].y * c.kg, gx[2*k+1].x * c.kg, gx[2*k+1].y * c.kg);
#pragma unroll
    for (int k = 0; k < 4; ++k) y4[4 + k] = make_float4(gy[2*k].x * c.kg, gy[2*k].y * c.kg, gy[2*k+1].x * c.kg, gy[2*k+1].y * c.kg);
    flags_s[p] = blur * c.cw > 0.1f ? 1 : 0;
}

// ---- shared-memory tile versions of the two hot sweeps (sweep_tile.cuh) ---------------------------------
template <int D>
__global__ void __launch_bounds__(128)
step_perceive_tile_kernel(GridView g, const float4* __restrict__ xv4, const float* __restrict__ A_s, StepConsts c,
                          float* __restrict__ Y_s, uint8_t* __restrict__ flags_s) {
    __shared__ TileSweepSmem<128, NF> sm;
    const int p0 = blockIdx.x * 128, p = p0 + threadIdx.x;
    const bool valid = p < g.n;
    const float4 xi = valid ? xv4[p] : make_float4(0.f, 0.f, 0.f, 0.f);
    float Ai[NF], gx[NF], gy[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) { Ai[f] = 0.f; gx[f] = 0.f; gy[f] = 0.f; }
    if (valid) load16(A_s + (size_t)p * NF, Ai);
    float blur = 0.f;
    tile_sweep<D, 128, NF, 128, 48>(g, xv4, A_s, sm, p0, c.h2, xi, valid, valid ? g.sorted_cell[p] : 0,
        [&](const float4& xj, const float* row) {
            float rx, ry;
            const float d2 = pair_d2<D>(xi, xj, rx, ry);
            float Aj[NF];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 t = *reinterpret_cast<const float4*>(row + 4 * k);
                Aj[4*k] = t.x; Aj[4*k+1] = t.y; Aj[4*k+2] = t.z; Aj[4*k+3] = t.w;
            }
            const float t = c.h2 - d2;
            const float alive = c.use_alpha ? (Aj[3] > 0.1f ? 1.f : 0.f) : 1.f;
            blur = fmaf(alive * xj.w, t * t * t, blur);
            if (d2 > 0.f) {
                const float d = sqrtf(d2);
                const float hd = c.h - d;
                const float m = 3.f * hd * hd / d * xj.w;
                const float wx = m * rx, wy = m * ry;
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    const float dA = Aj[f] - Ai[f];
                    gx[f] = fmaf(dA, wx, gx[f]);
                    gy[f] = fmaf(dA, wy, gy[f]);
                }
            }
        });
    if (!valid) return;
    float4* y4 = reinterpret_cast<float4*>(Y_s + (size_t)p * 2 * NF);
#pragma unroll
    for (int k = 0; k < 4; ++k) y4[k] = make_float4(gx[4*k] * c.kg, gx[4*k+1] * c.kg, gx[4*k+2] * c.kg, gx[4*k+3] * c.kg);
#pragma unroll
    for (int k = 0; k < 4; ++k) y4[4 + k] = make_float4(gy[4*k] * c.kg, gy[4*k+1] * c.kg, gy[4*k+2] * c.kg, gy[4*k+3] * c.kg);
    flags_s[p] = blur * c.cw > 0.1f ? 1 : 0;
}

template <int D>
__global__ void __launch_bounds__(128)
step_adjoint_tile_kernel(GridView g, const float4* __restrict__ xv4, const float* __restrict__ dY_s,
                         const float* __restrict__ dAdir_s, StepConsts c, float* __restrict__ dA_s) {
    __shared__ TileSweepSmem<128, 2 * NF> sm;
    const int p0 = blockIdx.x * 128, p = p0 + threadIdx.x;
    const bool valid = p < g.n;
    const float4 xi = valid ? xv4[p] : make_float4(0.f, 0.f, 0.f, 0.f);
    float acc[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] = 0.f;
    float sx = 0.f, sy = 0.f;
    tile_sweep<D, 128, 2 * NF, 128, 48>(g, xv4, dY_s, sm, p0, c.h2, xi, valid, valid ? g.sorted_cell[p] : 0,
        [&](const float4& xj, const float* row) {
            float rx, ry;
            const float d2 = pair_d2<D>(xi, xj, rx, ry);
            if (d2 > 0.f) {
                const float d = sqrtf(d2);
                const float hd = c.h - d;
                const float m = 3.f * hd * hd / d;
                const float wx = m * rx, wy = m * ry;
                sx = fmaf(wx, xj.w, sx); sy = fmaf(wy, xj.w, sy);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 a = *reinterpret_cast<const float4*>(row + 4 * k);
                    const float4 b = *reinterpret_cast<const float4*>(row + NF + 4 * k);
                    acc[4*k] = fmaf(wx, a.x, fmaf(wy, b.x, acc[4*k]));
                    acc[4*k+1] = fmaf(wx, a.y, fmaf(wy, b.y, acc[4*k+1]));
                    acc[4*k+2] = fmaf(wx, a.z, fmaf(wy, b.z, acc[4*k+2]));
                    acc[4*k+3] = fmaf(wx, a.w, fmaf(wy, b.w, acc[4*k+3]));
                }
            }
        });
    if (!valid) return;
    float Gx[NF], Gy[NF], dir[NF];
    load16(dY_s + (size_t)p * 2 * NF, Gx);
    load16(dY_s + (size_t)p * 2 * NF + NF, Gy);
    load16(dAdir_s + (size_t)p * NF, dir);
    float4* o4 = reinterpret_cast<float4*>(dA_s + (size_t)p * NF);
    float r[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) r[f] = dir[f] - c.kg * fmaf(xi.w, acc[f], fmaf(Gx[f], sx, Gy[f] * sy));
#pragma unroll
    for (int k = 0; k < 4; ++k) o4[k] = make_float4(r[4*k], r[4*k+1], r[4*k+2], r[4*k+3]);
}

// new life mask = blur(alive bits of the updated state) > 0.1; state *= prev & new   (nca.py:112-115)
template <int D>
__global__ void __launch_bounds__(256)
step_mask_kernel(GridView g, const float4* __restrict__ xv4, const uint8_t* __restrict__ alive_s, StepConsts c,
                 float* __restrict__ An_s, uint8_t* __restrict__ flags_s) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.n) return;
    const float4 xi = xv4[p];
    float blur = 0.f;
    for_each_candidate<D>(g, g.sorted_cell[p], [&](int q) {
        float rx, ry;
        const float4 xj = __ldg(&xv4[q]);
        const float d2 = pair_d2<D>(xi, xj, rx, ry);
        if (d2 < c.h2 && alive_s[q]) {
            const float t = c.h2 - d2;
            blur = fmaf(xj.w, t * t * t, blur);
        }
    });
    const bool new_mask = blur * c.cw > 0.1f;
    uint8_t fl = flags_s[p];
    if (new_mask) fl |= 2;
    flags_s[p] = fl;
    if (!((fl & 1) && new_mask)) {
        float4* a4 = reinterpret_cast<float4*>(An_s + (size_t)p * NF);
#pragma unroll
        for (int k = 0; k < 4; ++k) { float4 t = a4[k]; a4[k] = make_float4(t.x * 0.f, t.y * 0.f, t.z * 0.f, t.w * 0.f); }
    }
}

// Data gradient of everything after the perception sweep, one particle per thread.
//   in : dA_next (grad of the masked output), saved A, Y, O, flags
//   out: dO (grad wrt MLP output, for the weight-gradient stage), dY (grad wrt scaled perception),
//        dAdir (grad reaching A directly: gate path, fire bypass, MLP input columns 0..15)
__global__ void __launch_bounds__(kStepThreads, 1)
step_bwd_kernel(int n_total, const float* __restrict__ A_s, const float* __restrict__ Y_s,
                const float* __restrict__ O_s, const uint8_t* __restrict__ flags_s,
                const float* __restrict__ dAn_s, sphnca_step_params prm, StepConsts c,
                float* __restrict__ dO_s, float* __restrict__ dY_s, float* __restrict__ dAdir_s,
                float* __restrict__ Ht, float* __restrict__ dHt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MlpSmem& s = *reinterpret_cast<MlpSmem*>(smem_raw);
    stage_weights(s, prm, c.n_out);
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_total) return;

    float y[NIN], dnA[NF], dO[NO_PAD], dAdir[NF];
    load16(A_s + (size_t)p * NF, *reinterpret_cast<float(*)[NF]>(&y[0]));
    load16(Y_s + (size_t)p * 2 * NF, *reinterpret_cast<float(*)[NF]>(&y[NF]));
    load16(Y_s + (size_t)p * 2 * NF + NF, *reinterpret_cast<float(*)[NF]>(&y[2 * NF]));
    load16(dAn_s + (size_t)p * NF, dnA);
    const uint8_t fl = flags_s[p];
    const bool living = (fl & 1) && (fl & 2), fire = fl & 4;
#pragma unroll
    for (int f = 0; f < NF; ++f) dnA[f] = living ? dnA[f] : 0.f;
#pragma unroll
    for (int o = 0; o < NO_PAD; ++o) dO[o] = 0.f;
    if (fire) {
        const float4* o4 = reinterpret_cast<const float4*>(O_s + (size_t)p * NO_PAD);
        float out[NO_PAD];
#pragma unroll
        for (int k = 0; k < NO_PAD / 4; ++k) { float4 t = __ldg(o4 + k); out[4*k] = t.x; out[4*k+1] = t.y; out[4*k+2] = t.z; out[4*k+3] = t.w; }
        if (c.rule == 0) {
            const float mult = sigmoidf_(out[2 * NF]);
            float dmult = 0.f;
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const float gate = sigmoidf_(out[f]);
                const float delta = tanhf(out[NF + f]);
                dO[f] = dnA[f] * y[f] * gate * (1.f - gate);
                dO[NF + f] = dnA[f] * mult * (1.f - delta * delta);
                dmult = fmaf(dnA[f], delta, dmult);
                dAdir[f] = dnA[f] * gate;
            }
            dO[2 * NF] = dmult * mult * (1.f - mult);
        } else {
#pragma unroll
            for (int f = 0; f < NF; ++f) { dO[f] = dnA[f] * c.orig_rate; dAdir[f] = dnA[f]; }
        }
    } else {
#pragma unroll
        for (int f = 0; f < NF; ++f) dAdir[f] = dnA[f];
    }
    float4* do4 = reinterpret_cast<float4*>(dO_s + (size_t)p * NO_PAD);
#pragma unroll
    for (int k = 0; k < NO_PAD / 4; ++k) do4[k] = make_float4(dO[4*k], dO[4*k+1], dO[4*k+2], dO[4*k+3]);

    // dy = W1^T ( relu'(pre) * (W2^T dO) ), hidden layer recomputed.  Warp-uniform skip when no
    // lane fired (dO == 0 everywhere in the warp).
    float dy[NIN];
#pragma unroll
    for (int k = 0; k < NIN; ++k) dy[k] = 0.f;
    if (__any_sync(__activemask(), fire)) {
#pragma unroll 1
        for (int n = 0; n < NH; ++n) {
            const float pre = hidden_pre(s, n, y);
            const float4* w2 = reinterpret_cast<const float4*>(&s.W2t[n][0]);
            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
            for (int k = 0; k < NO_PAD / 4; ++k) {
                const float4 t = w2[k];
                d0 = fmaf(t.x, dO[4*k], d0); d1 = fmaf(t.y, dO[4*k+1], d1);
                d2 = fmaf(t.z, dO[4*k+2], d2); d3 = fmaf(t.w, dO[4*k+3], d3);
            }
            const float dh = pre > 0.f ? (d0 + d1) + (d2 + d3) : 0.f;
            if (Ht) {   // tile-blocked [p/32][256][32]: a warp writes 128 contiguous bytes per hidden unit
                const size_t o = ((size_t)(p >> 5) * NH + n) * 32 + (p & 31);
                Ht[o] = fmaxf(pre, 0.f);
                dHt[o] = dh;
            }
            const float4* w1 = reinterpret_cast<const float4*>(&s.W1[n][0]);
#pragma unroll
            for (int k = 0; k < NIN / 4; ++k) {
                const float4 t = w1[k];
                dy[4*k] = fmaf(t.x, dh, dy[4*k]); dy[4*k+1] = fmaf(t.y, dh, dy[4*k+1]);
                dy[4*k+2] = fmaf(t.z, dh, dy[4*k+2]); dy[4*k+3] = fmaf(t.w, dh, dy[4*k+3]);
            }
        }
    } else if (Ht) {
        // no lane of this warp fired: dO == 0, so dH == 0 and H is multiplied by zero downstream
#pragma unroll 4
        for (int n = 0; n < NH; ++n) {
            const size_t o = ((size_t)(p >> 5) * NH + n) * 32 + (p & 31);
            Ht[o] = 0.f; dHt[o] = 0.f;
        }
    }
    float4* d4 = reinterpret_cast<float4*>(dAdir_s + (size_t)p * NF);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        d4[k] = make_float4(dAdir[4*k] + dy[4*k], dAdir[4*k+1] + dy[4*k+1], dAdir[4*k+2] + dy[4*k+2], dAdir[4*k+3] + dy[4*k+3]);
    float4* g4 = reinterpret_cast<float4*>(dY_s + (size_t)p * 2 * NF);
#pragma unroll
    for (int k = 0; k < 8; ++k) g4[k] = make_float4(dy[NF+4*k], dy[NF+4*k+1], dy[NF+4*k+2], dy[NF+4*k+3]);
}

// Gather adjoint of the perception sweep (same algebra as gradient_bwd_kernel, x/y only):
//   dA_i[f] = dAdir_i[f] - kg ( v_i sum_j m_ij (r_x Gx_j[f] + r_y Gy_j[f]) + Gx_i[f] s_x + Gy_i[f] s_y ),
//   s = sum_j m_ij v_j r,  G = dL/dY.
template <int D>
__global__ void __launch_bounds__(128)
step_adjoint_kernel(GridView g, const float4* __restrict__ xv4, const float* __restrict__ dY_s,
                    const float* __restrict__ dAdir_s, StepConsts c, float* __restrict__ dA_s) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.n) return;
    const float4 xi = xv4[p];
    float2 acc[NF / 2];                                   // channel pairs (FFMA2)
#pragma unroll
    for (int f = 0; f < NF / 2; ++f) acc[f] = make_float2(0.f, 0.f);
    float sx = 0.f, sy = 0.f;
    for_each_candidate<D>(g, g.sorted_cell[p], [&](int q) {
        float rx, ry;
        const float4 xj = __ldg(&xv4[q]);
        const float d2 = pair_d2<D>(xi, xj, rx, ry);
        if (d2 < c.h2 && d2 > 0.f) {
            const float inv = rsqrtf(d2);
            const float hd = c.h - d2 * inv;
            const float m = 3.f * hd * hd * inv;
            const float wxs = m * rx, wys = m * ry;
            sx = fmaf(wxs, xj.w, sx); sy = fmaf(wys, xj.w, sy);
            const float2 wx = make_float2(wxs, wxs), wy = make_float2(wys, wys);
            const float4* g4 = reinterpret_cast<const float4*>(dY_s + (size_t)q * 2 * NF);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 a = __ldg(g4 + k), b = __ldg(g4 + 4 + k);
                acc[2*k] = ffma2(wx, make_float2(a.x, a.y), ffma2(wy, make_float2(b.x, b.y), acc[2*k]));
                acc[2*k+1] = ffma2(wx, make_float2(a.z, a.w), ffma2(wy, make_float2(b.z, b.w), acc[2*k+1]));
            }
        }
    });
    float Gx[NF], Gy[NF], dir[NF];
    load16(dY_s + (size_t)p * 2 * NF, Gx);
    load16(dY_s + (size_t)p * 2 * NF + NF, Gy);
    load16(dAdir_s + (size_t)p * NF, dir);
    float4* o4 = reinterpret_cast<float4*>(dA_s + (size_t)p * NF);
    float r[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        const float a = (f & 1) ? acc[f >> 1].y : acc[f >> 1].x;
        r[f] = dir[f] - c.kg * fmaf(xi.w, a, fmaf(Gx[f], sx, Gy[f] * sy));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) o4[k] = make_float4(r[4*k], r[4*k+1], r[4*k+2], r[4*k+3]);
}

}  // namespace sphnca

using namespace sphnca;

static int g_sweep_path = 0;

extern "C" {

}  // extern "C"

namespace sphnca {
// ---- life mask without an alpha channel ------------------------------------------------------------------------
// use_alpha == 0: cell_activity is all-ones (reference nca.py:17-21), so the new life mask is the blur of the same
// constant field as the previous one -- the value the perception sweep already wrote into bit 0, bit for bit.  No sweep:
// bit1 = bit0, dead rows are multiplied by 0 exactly like step_mask_kernel does.
__global__ void __launch_bounds__(256) step_mask_static_kernel(int64_t n, float* __restrict__ An_s, uint8_t* __restrict__ flags_s) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint8_t fl = flags_s[p];
    if (fl & 1) { flags_s[p] = fl | 2; return; }
    float4* a4 = reinterpret_cast<float4*>(An_s + (size_t)p * NF);
#pragma unroll
    for (int k = 0; k < 4; ++k) { float4 t = a4[k]; a4[k] = make_float4(t.x * 0.f, t.y * 0.f, t.z * 0.f, t.w * 0.f); }
}

// ---- perception sweep with block-staged candidates (sweep_stage.cuh) --------------------------------------------
// Opt-in (sphnca_set_sweep_path(2)).  Measured on B200, 524,288 particles: padded 2-D (BASELINE configs[1]) 0.266 ms
// vs 0.300 ms for the gather kernel; 3-D clouds never fit the staging buffer (27 cells x 7.6 particles x 9 rows),
// every tile falls back and pays the planning + the lower occupancy (0.475 vs 0.332 ms) -- hence not the default.
// The two-pass variant (PERC_SLOTS > 0, padded rows, 256-particle tiles) executes 9 % fewer instructions at 25 instead
// of 18 active lanes but issues on 31 % instead of 53 % of the cycles: 0.387 ms.
constexpr int PERC_T = 128;        // particles per tile (~12 cells of a row in the padded 2-D configuration)
constexpr int PERC_CAP = 552;      // candidates per tile: 552 x (16 + 64) B = 44 KB of shared memory -> 5 CTAs per SM
constexpr int PERC_SLOTS = 0;      // 0 = single pass; > 0 = hit-list slots per thread of the two-pass variant
template <int D>
__global__ void __launch_bounds__(PERC_T, 5)
step_perceive_staged_kernel(GridView g, const float4* __restrict__ xv4, const float* __restrict__ A_s, StepConsts c,
                            float* __restrict__ Y_s, uint8_t* __restrict__ flags_s) {
    extern __shared__ __align__(128) unsigned char stage_raw[];
    using S = StageSmem<D, PERC_T, NF, PERC_CAP, PERC_SLOTS>;
    S& sm = *reinterpret_cast<S*>(stage_raw);
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < g.n;
    const int my_cell = valid ? g.sorted_cell[p] : 0;
    int k = 0;
    const bool staged = stage_tile(g, xv4, A_s, sm, p, valid, my_cell, k);
    if (!valid) return;
    const float4 xi = xv4[p];
    float2 Ai[NF / 2], gx[NF / 2], gy[NF / 2];
    {
        const float4* a4 = reinterpret_cast<const float4*>(A_s + (size_t)p * NF);
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float4 t = __ldg(a4 + q); Ai[2*q] = make_float2(t.x, t.y); Ai[2*q+1] = make_float2(t.z, t.w); }
    }
#pragma unroll
    for (int f = 0; f < NF / 2; ++f) { gx[f] = make_float2(0.f, 0.f); gy[f] = make_float2(0.f, 0.f); }
    float blur = 0.f;
    // the pair body of step_perceive_kernel, unchanged; `row4(q)` yields float4 q of the neighbour's 16 channels
    auto pair = [&](const float4 xj, auto&& row4) {
        float rx, ry;
        const float d2 = pair_d2<D>(xi, xj, rx, ry);
        if (d2 < c.h2) {
            float2 Aj[NF / 2];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const float4 t = row4(q); Aj[2*q] = make_float2(t.x, t.y); Aj[2*q+1] = make_float2(t.z, t.w); }
            const float t = c.h2 - d2;
            const float alive = c.use_alpha ? (Aj[1].y > 0.1f ? 1.f : 0.f) : 1.f;
            blur = fmaf(alive * xj.w, t * t * t, blur);
            if (d2 > 0.f) {
                const float inv = rsqrtf(d2);
                const float hd = c.h - d2 * inv;
                const float m = 3.f * hd * hd * inv * xj.w;
                const float2 wx = make_float2(m * rx, m * rx), wy = make_float2(m * ry, m * ry);
#pragma unroll
                for (int f = 0; f < NF / 2; ++f) {
                    const float2 dA = fsub2(Aj[f], Ai[f]);
                    gx[f] = ffma2(dA, wx, gx[f]);
                    gy[f] = ffma2(dA, wy, gy[f]);
                }
            }
        }
    };
    if (staged) {
        if (PERC_SLOTS > 0)
            staged_two_pass(g, sm, my_cell, k,
                [&](const float4 xj) { float rx, ry; return pair_d2<D>(xi, xj, rx, ry) < c.h2; },
                [&](int j) { pair(sm.x[j], [&](int q) { return sm.row[j * S::RP4 + q]; }); });
        else
            for_each_staged(g, sm, my_cell, k, [&](int j) { pair(sm.x[j], [&](int q) { return sm.row[j * S::RP4 + q]; }); });
    } else {
        for_each_candidate<D>(g, my_cell, [&](int q_) {
            const float4* a4 = reinterpret_cast<const float4*>(A_s + (size_t)q_ * NF);
            pair(__ldg(&xv4[q_]), [&](int q) { return __ldg(a4 + q); });
        });
    }
    float4* y4 = reinterpret_cast<float4*>(Y_s + (size_t)p * 2 * NF);
#pragma unroll
    for (int q = 0; q < 4; ++q) y4[q] = make_float4(gx[2*q].x * c.kg, gx[2*q].y * c.kg, gx[2*q+1].x * c.kg, gx[2*q+1].y * c.kg);
#pragma unroll
    for (int q = 0; q < 4; ++q) y4[4 + q] = make_float4(gy[2*q].x * c.kg, gy[2*q].y * c.kg, gy[2*q+1].x * c.kg, gy[2*q+1].y * c.kg);
    flags_s[p] = blur * c.cw > 0.1f ? 1 : 0;
}

// counts[0] += 1 per tile, counts[1] += 1 per tile that does not fit the staging buffer of step_perceive_staged_kernel
template <int D>
__global__ void __launch_bounds__(PERC_T) stage_fit_kernel(GridView g, int32_t* __restrict__ counts) {
    extern __shared__ __align__(128) unsigned char stage_raw[];
    using S = StageSmem<D, PERC_T, NF, PERC_CAP, PERC_SLOTS>;
    S& sm = *reinterpret_cast<S*>(stage_raw);
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < g.n;
    const int my_cell = valid ? g.sorted_cell[p] : 0;
    int k = 0, nseg = 0;
    const bool fits = plan_tile<D, PERC_T, NF, PERC_CAP, PERC_SLOTS, 4>(g, nullptr, nullptr, sm, p, valid, my_cell, k, nseg);
    if (threadIdx.x == 0) { atomicAdd(&counts[0], 1); if (!fits) atomicAdd(&counts[1], 1); }
}
}  // namespace sphnca

extern "C" {
using namespace sphnca;

int32_t sphnca_get_mlp_path(void) { return mlp_path(); }

int32_t sphnca_set_sweep_path(int32_t path) {
    if (path < 0 || path > 2) { set_error("sweep path must be 0 (per-thread gathers), 1 (shared-memory tile engine) or 2 (TMA-staged tiles)"); return SPHNCA_ERR_INVALID; }
    g_sweep_path = path;
    return SPHNCA_OK;
}

static int launch_perceive_staged(const GridView& gv, const StepConsts& c, const float* xv4, const float* A_s, float* Y_s,
                                  uint8_t* flags_s, cudaStream_t stream) {
    const int sblocks = (int)((gv.n + PERC_T - 1) / PERC_T);
    if (gv.dim == 3) {
        const size_t smem = sizeof(StageSmem<3, PERC_T, NF, PERC_CAP, PERC_SLOTS>);
        SPHNCA_CUDA(cudaFuncSetAttribute(step_perceive_staged_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SPHNCA_LAUNCH(step_perceive_staged_kernel<3>, sblocks, PERC_T, smem, stream, gv, (const float4*)xv4, A_s, c, Y_s, flags_s);
    } else {
        const size_t smem = sizeof(StageSmem<2, PERC_T, NF, PERC_CAP, PERC_SLOTS>);
        SPHNCA_CUDA(cudaFuncSetAttribute(step_perceive_staged_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SPHNCA_LAUNCH(step_perceive_staged_kernel<2>, sblocks, PERC_T, smem, stream, gv, (const float4*)xv4, A_s, c, Y_s, flags_s);
    }
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_step_perceive_staged(const sphnca_grid* g, const float* xv4, const float* A_s, const sphnca_step_params* p,
                                    float* Y_s, uint8_t* flags_s, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv); if (rc) return rc; }
    StepConsts c; { int rc = make_consts(g, p, &c); if (rc) return rc; }
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(xv4 && A_s && Y_s && flags_s, "NULL argument");
    return launch_perceive_staged(gv, c, xv4, A_s, Y_s, flags_s, as_stream(stream_));
}

int32_t sphnca_stage_fit(const sphnca_grid* g, int32_t* counts, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv); if (rc) return rc; }
    SPHNCA_CHECK_ARG(counts, "NULL argument");
    cudaStream_t stream = as_stream(stream_);
    SPHNCA_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(int32_t), stream));
    if (gv.n == 0) return SPHNCA_OK;
    const int sblocks = (int)((gv.n + PERC_T - 1) / PERC_T);
    if (gv.dim == 3) {
        const size_t smem = sizeof(StageSmem<3, PERC_T, NF, PERC_CAP, PERC_SLOTS>);
        SPHNCA_CUDA(cudaFuncSetAttribute(stage_fit_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SPHNCA_LAUNCH(stage_fit_kernel<3>, sblocks, PERC_T, smem, stream, gv, counts);
    } else {
        const size_t smem = sizeof(StageSmem<2, PERC_T, NF, PERC_CAP, PERC_SLOTS>);
        SPHNCA_CUDA(cudaFuncSetAttribute(stage_fit_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SPHNCA_LAUNCH(stage_fit_kernel<2>, sblocks, PERC_T, smem, stream, gv, counts);
    }
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_step_perceive(const sphnca_grid* g, const float* xv4, const float* A_s, const sphnca_step_params* p,
                             float* Y_s, uint8_t* flags_s, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv); if (rc) return rc; }
    StepConsts c; { int rc = make_consts(g, p, &c); if (rc) return rc; }
    cudaStream_t stream = as_stream(stream_);
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(xv4 && A_s && Y_s && flags_s, "NULL argument");
    const int pblocks = (int)((gv.n + 127) / 128);
    if (g_sweep_path == 2) {
        return launch_perceive_staged(gv, c, xv4, A_s, Y_s, flags_s, stream);
    } else if (g_sweep_path == 1) {
        if (gv.dim == 3) SPHNCA_LAUNCH(step_perceive_tile_kernel<3>, pblocks, 128, 0, stream, gv, (const float4*)xv4, A_s, c, Y_s, flags_s);
        else SPHNCA_LAUNCH(step_perceive_tile_kernel<2>, pblocks, 128, 0, stream, gv, (const float4*)xv4, A_s, c, Y_s, flags_s);
    } else {
        if (gv.dim == 3) SPHNCA_LAUNCH(step_perceive_kernel<3>, pblocks, 128, 0, stream, gv, (const float4*)xv4, A_s, c, Y_s, flags_s);
        else SPHNCA_LAUNCH(step_perceive_kernel<2>, pblocks, 128, 0, stream, gv, (const float4*)xv4, A_s, c, Y_s, flags_s);
    }
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_step_mlp(int64_t N, const float* A_s, const float* Y_s, const uint8_t* fire_s, const sphnca_step_params* p,
                        float* A_new_s, uint8_t* alive_s, float* O_s, uint8_t* flags_s, uint32_t* relu_s, void* stream_) {
    SPHNCA_CHECK_ARG(N >= 0 && N < (int64_t(1) << 31), "bad N");
    sphnca_grid fake{}; fake.dim = 3; fake.h = 1.0;      // geometry constants are unused by the MLP half
    StepConsts c; { int rc = make_consts(&fake, p, &c); if (rc) return rc; }
    if (N == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(A_s && Y_s && fire_s && A_new_s && alive_s && O_s && flags_s && relu_s, "NULL argument");
    return launch_step_mlp_tc(N, c, p, A_s, Y_s, fire_s, A_new_s, alive_s, O_s, flags_s, relu_s, as_stream(stream_));
}

int32_t sphnca_step_mlp_list(int64_t N, const int32_t* rows, const int32_t* count, const float* A_s, const float* Y_s,
                             const sphnca_step_params* p, float* A_new_s, uint8_t* alive_s, float* O_c, uint8_t* flags_s,
                             uint32_t* relu_c, float* A_c, float* Y_c, int32_t inputs_by_list, void* stream_) {
    SPHNCA_CHECK_ARG(N >= 0 && N < (int64_t(1) << 31), "bad N");
    sphnca_grid fake{}; fake.dim = 3; fake.h = 1.0;      // geometry constants are unused by the MLP half
    StepConsts c; { int rc = make_consts(&fake, p, &c); if (rc) return rc; }
    if (N == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(rows && count && A_s && Y_s && A_new_s && alive_s && O_c && flags_s && relu_c, "NULL argument");
    SPHNCA_CHECK_ARG((A_c == nullptr) == (Y_c == nullptr), "A_c and Y_c must be given together");
    SPHNCA_CHECK_ARG(!((inputs_by_list & 1) && A_c), "inputs_by_list: A_s / Y_s are the compact arrays already, no copies are made");
    // inputs_by_list bit 0: A_s / Y_s are compact arrays; bit 1: apply the static life mask here (flags from sphnca_fire_list_draw)
    SPHNCA_CHECK_ARG(!(inputs_by_list & 2) || !p->use_alpha, "the static life mask exists without an alpha channel only");
    return launch_step_mlp_tc(N, c, p, A_s, Y_s, nullptr, A_new_s, alive_s, O_c, flags_s, relu_c, as_stream(stream_), rows, count, A_c, Y_c,
                              inputs_by_list & 1, (inputs_by_list >> 1) & 1);
}

int32_t sphnca_step_fwd(const sphnca_grid* g, const float* xv4, const float* A_s, const uint8_t* fire_s,
                        const sphnca_step_params* p, float* A_new_s, uint8_t* alive_s, float* Y_s, float* O_s,
                        uint8_t* flags_s, uint32_t* relu_s, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv); if (rc) return rc; }
    StepConsts c; { int rc = make_consts(g, p, &c); if (rc) return rc; }
    cudaStream_t stream = as_stream(stream_);
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(xv4 && A_s && fire_s && A_new_s && alive_s && Y_s && O_s && flags_s && relu_s, "NULL argument");
    if (mlp_path() == 2)      // fused tcgen05 kernel (sweep + MLP in one persistent kernel)
        return launch_step_fwd_tc(gv, c, p, xv4, A_s, fire_s, A_new_s, alive_s, Y_s, O_s, flags_s, relu_s, stream);
    if (mlp_path() == 1) {    // split: full-L1 sweep kernel, then the tcgen05 MLP kernel (default)
        int rc = sphnca_step_perceive(g, xv4, A_s, p, Y_s, flags_s, stream_);
        if (rc) return rc;
        return sphnca_step_mlp(gv.n, A_s, Y_s, fire_s, p, A_new_s, alive_s, O_s, flags_s, relu_s, stream_);
    }
    const int blocks = (int)((gv.n + kStepThreads - 1) / kStepThreads);
    const size_t smem = sizeof(MlpSmem);
    if (gv.dim == 3) {
        SPHNCA_CUDA(cudaFuncSetAttribute(step_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SPHNCA_LAUNCH(step_fwd_kernel<3>, blocks, kStepThreads, smem, stream, gv, (const float4*)xv4, A_s, fire_s, *p, c, A_new_s, Y_s, O_s, flags_s, alive_s, relu_s);
    } else {
        SPHNCA_CUDA(cudaFuncSetAttribute(step_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SPHNCA_LAUNCH(step_fwd_kernel<2>, blocks, kStepThreads, smem, stream, gv, (const float4*)xv4, A_s, fire_s, *p, c, A_new_s, Y_s, O_s, flags_s, alive_s, relu_s);
    }
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_step_mask(const sphnca_grid* g, const float* xv4, const uint8_t* alive_s, float* A_new_s,
                         uint8_t* flags_s, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv); if (rc) return rc; }
    cudaStream_t stream = as_stream(stream_);
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(xv4 && alive_s && A_new_s && flags_s, "NULL argument");
    StepConsts c{};
    c.h = (float)g->h; c.h2 = (float)(g->h * g->h); c.cw = (float)poly6_norm(g->dim, g->h);
    const int blocks = (int)((gv.n + 255) / 256);
    if (gv.dim == 3) SPHNCA_LAUNCH(step_mask_kernel<3>, blocks, 256, 0, stream, gv, (const float4*)xv4, alive_s, c, A_new_s, flags_s);
    else SPHNCA_LAUNCH(step_mask_kernel<2>, blocks, 256, 0, stream, gv, (const float4*)xv4, alive_s, c, A_new_s, flags_s);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_step_mask_static(int64_t N, float* A_new_s, uint8_t* flags_s, void* stream_) {
    SPHNCA_CHECK_ARG(N >= 0 && N < (int64_t(1) << 31), "bad N");
    if (N == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(A_new_s && flags_s, "NULL argument");
    SPHNCA_LAUNCH(step_mask_static_kernel, (int)((N + 255) / 256), 256, 0, as_stream(stream_), N, A_new_s, flags_s);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_step_bwd(int64_t N, const float* A_s, const float* Y_s, const float* O_s, const uint8_t* flags_s,
                        const uint32_t* relu_s, const float* dA_next_s, const sphnca_step_params* p, float* dO_s,
                        float* dY_s, float* dAdir_s, float* Ht, float* dHt, void* stream_) {
    SPHNCA_CHECK_ARG(N >= 0 && N < (int64_t(1) << 31), "bad N");
    sphnca_grid fake{}; fake.dim = 3; fake.h = 1.0;      // only the rule / rates of the constants are used here
    StepConsts c; { int rc = make_consts(&fake, p, &c); if (rc) return rc; }
    cudaStream_t stream = as_stream(stream_);
    if (N == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(A_s && Y_s && O_s && flags_s && dA_next_s && dO_s && dY_s && dAdir_s, "NULL argument");
    SPHNCA_CHECK_ARG((Ht == nullptr) == (dHt == nullptr), "Ht and dHt must be given together");
    if (mlp_path() >= 1 && Ht == nullptr) {
        SPHNCA_CHECK_ARG(relu_s != nullptr, "relu_s is required by the tensor-core data-gradient path");
        return launch_step_bwd_tc(N, c, p, A_s, O_s, flags_s, relu_s, dA_next_s, dO_s, dY_s, dAdir_s, stream);
    }
    const size_t smem = sizeof(MlpSmem);
    SPHNCA_CUDA(cudaFuncSetAttribute(step_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SPHNCA_LAUNCH(step_bwd_kernel, (int)((N + kStepThreads - 1) / kStepThreads), kStepThreads, smem, stream,
                  (int)N, A_s, Y_s, O_s, flags_s, dA_next_s, *p, c, dO_s, dY_s, dAdir_s, Ht, dHt);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_step_bwd_list(int64_t N, const int32_t* rows, const int32_t* count, const float* A_c, const float* O_c,
                             const uint8_t* flags_s, const uint32_t* relu_c, const float* dA_next_s,
                             const sphnca_step_params* p, float* dO_c, float* dY_s, float* dAdir_s, void* stream_) {
    SPHNCA_CHECK_ARG(N >= 0 && N < (int64_t(1) << 31), "bad N");
    sphnca_grid fake{}; fake.dim = 3; fake.h = 1.0;
    StepConsts c; { int rc = make_consts(&fake, p, &c); if (rc) return rc; }
    if (N == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(rows && count && A_c && O_c && flags_s && relu_c && dA_next_s && dO_c && dY_s && dAdir_s, "NULL argument");
    return launch_step_bwd_tc(N, c, p, A_c, O_c, flags_s, relu_c, dA_next_s, dO_c, dY_s, dAdir_s, as_stream(stream_), rows, count);
}

int32_t sphnca_step_adjoint(const sphnca_grid* g, const float* xv4, const float* dY_s, const float* dAdir_s,
                            const sphnca_step_params* p, float* dA_s, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv); if (rc) return rc; }
    StepConsts c; { int rc = make_consts(g, p, &c); if (rc) return rc; }
    cudaStream_t stream = as_stream(stream_);
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(xv4 && dY_s && dAdir_s && dA_s, "NULL argument");
    const int blocks = (int)((gv.n + 127) / 128);
    if (g_sweep_path == 1) {
        if (gv.dim == 3) SPHNCA_LAUNCH(step_adjoint_tile_kernel<3>, blocks, 128, 0, stream, gv, (const float4*)xv4, dY_s, dAdir_s, c, dA_s);
        else SPHNCA_LAUNCH(step_adjoint_tile_kernel<2>, blocks, 128, 0, stream, gv, (const float4*)xv4, dY_s, dAdir_s, c, dA_s);
    } else if (gv.dim == 3) SPHNCA_LAUNCH(step_adjoint_kernel<3>, blocks, 128, 0, stream, gv, (const float4*)xv4, dY_s, dAdir_s, c, dA_s);
    else SPHNCA_LAUNCH(step_adjoint_kernel<2>, blocks, 128, 0, stream, gv, (const float4*)xv4, dY_s, dAdir_s, c, dA_s);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

}  // extern "C"
