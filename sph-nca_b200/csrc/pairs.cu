// Pair lists: the geometry of a rollout is fixed (positions and volumes do not change over the K steps of
// reference nca.py:141-145), so everything the two hot sweeps recompute per step and per candidate -- the 3^D-cell
// walk, the in-radius test (only ~35 % of the candidates of a padded 2-D cloud and ~15 % of a 3-D cloud pass it) and
// the kernel weights -- is done ONCE per hash grid and stored per in-radius pair:
//
//   J   int32   sorted position of the neighbour
//   WP  float2  perception weights  m v_j (x_j - x_i)_{x,y},  m = 3 (h - d)^2 / d   (0 for d = 0)
//   WB  float2  {v_j, (h^2 - d^2)^3}  for the life-mask blur
//   WA  float2  adjoint weights     m (x_j - x_i)_{x,y}
//   SXY float2  per particle: sum_j WA v_j (the centre term of the adjoint)
//
// Layout: per warp of 32 consecutive cell-sorted particles, [slot][lane]; a warp's slot count is the maximum
// neighbour count of its lanes, shorter lists are padded with zero-weight self pairs.  The sweeps then run a
// warp-uniform loop (no divergence, no tests) of coalesced 4/8-byte list loads + one row gather + the channel FMAs.
// Entries are appended in the candidate order of for_each_candidate and hold the very floats the gather kernels
// compute, so the sums are bit-identical to step_perceive_kernel / step_adjoint_kernel.
// Cost: ~32 B per pair (about 1.3 KB per particle), one count + one fill sweep per grid.
#include "step_common.cuh"

namespace sphnca {

template <int D>
__global__ void __launch_bounds__(128)
pairs_count_kernel(GridView g, const float4* __restrict__ xv4, float h2, int32_t* __restrict__ warp_slots) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    if (p < g.n) {
        const float4 xi = xv4[p];
        for_each_candidate<D>(g, g.sorted_cell[p], [&](int q) {
            float rx, ry;
            if (pair_d2<D>(xi, __ldg(&xv4[q]), rx, ry) < h2) ++cnt;
        });
    }
    const int mx = __reduce_max_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && p < g.n) warp_slots[p >> 5] = mx;
}

template <int D>
__global__ void __launch_bounds__(128)
pairs_fill_kernel(GridView g, const float4* __restrict__ xv4, float h, float h2, const int32_t* __restrict__ warp_slots,
                  const int64_t* __restrict__ warp_base, int32_t* __restrict__ J, float2* __restrict__ WP,
                  float2* __restrict__ WB, float2* __restrict__ WA, float2* __restrict__ SXY, float cw,
                  uint8_t* __restrict__ FLG) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t w = p >> 5;
    if (w * 32 >= g.n) return;
    const int lane = threadIdx.x & 31;
    const int slots = warp_slots[w];
    int64_t e = warp_base[w] + lane;
    int cnt = 0;
    const int64_t self = p < g.n ? p : g.n - 1;
    if (p < g.n) {
        const float4 xi = xv4[p];
        float sx = 0.f, sy = 0.f, blur1 = 0.f;
        for_each_candidate<D>(g, g.sorted_cell[p], [&](int q) {
            float rx, ry;
            const float4 xj = __ldg(&xv4[q]);
            const float d2 = pair_d2<D>(xi, xj, rx, ry);
            if (d2 < h2) {
                const float t = h2 - d2;
                blur1 = fmaf(1.f * xj.w, t * t * t, blur1);           // the life-mask blur of an all-alive field (no alpha channel)
                float2 wp = make_float2(0.f, 0.f), wa = make_float2(0.f, 0.f);
                if (d2 > 0.f) {
                    // the expressions of step_perceive_kernel / step_adjoint_kernel, operation for operation
                    const float inv = rsqrtf(d2);
                    const float hd = h - d2 * inv;
                    const float mp = 3.f * hd * hd * inv * xj.w;
                    wp = make_float2(mp * rx, mp * ry);
                    const float ma = 3.f * hd * hd * inv;
                    wa = make_float2(ma * rx, ma * ry);
                    sx = fmaf(wa.x, xj.w, sx); sy = fmaf(wa.y, xj.w, sy);
                }
                J[e] = q; WP[e] = wp; WB[e] = make_float2(xj.w, t * t * t); WA[e] = wa;
                e += 32; ++cnt;
            }
        });
        SXY[p] = make_float2(sx, sy);
        FLG[p] = blur1 * cw > 0.1f ? 1 : 0;
    }
    for (; cnt < slots; ++cnt, e += 32) {                 // zero-weight self pairs up to the warp's slot count
        J[e] = (int32_t)self; WP[e] = make_float2(0.f, 0.f); WB[e] = make_float2(0.f, 0.f); WA[e] = make_float2(0.f, 0.f);
    }
}

// Perception sweep + previous life mask from the pair lists (same outputs as step_perceive_kernel).
// One lane per particle: the three list arrays are read fully coalesced (128/256 bytes per warp and slot); the 64-byte
// neighbour row is gathered with four 16-byte loads per lane.  (Four lanes per particle, as in the adjoint below, and a
// shuffle-distributed hybrid were measured at 0.24 / 0.23 ms against 0.21 ms for this form.)
// ALPHA = false (no alpha channel): the life mask does not depend on the state; it was evaluated once by the fill kernel
// (FLG) and the blur weights WB are not read at all.
template <bool ALPHA>
__global__ void __launch_bounds__(128)
step_perceive_pairs_kernel(int64_t n, const int32_t* __restrict__ warp_slots, const int64_t* __restrict__ warp_base,
                           const int32_t* __restrict__ J, const float2* __restrict__ WP, const float2* __restrict__ WB,
                           const uint8_t* __restrict__ FLG, const float* __restrict__ A_s, StepConsts c,
                           float* __restrict__ Y_s, uint8_t* __restrict__ flags_s) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t w = p >> 5;
    if (w * 32 >= n) return;
    const int64_t pi = p < n ? p : n - 1;
    const int slots = warp_slots[w];
    const int64_t e0 = warp_base[w] + (threadIdx.x & 31);
    float2 Ai[NF / 2], gx[NF / 2], gy[NF / 2];
    {
        const float4* a4 = reinterpret_cast<const float4*>(A_s + (size_t)pi * NF);
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float4 t = __ldg(a4 + k); Ai[2*k] = make_float2(t.x, t.y); Ai[2*k+1] = make_float2(t.z, t.w); }
    }
#pragma unroll
    for (int f = 0; f < NF / 2; ++f) { gx[f] = make_float2(0.f, 0.f); gy[f] = make_float2(0.f, 0.f); }
    float blur = 0.f;
    // list entries one slot ahead of the arithmetic (streaming loads: read once per sweep)
    int j = 0; float2 wp = make_float2(0.f, 0.f), wb = make_float2(0.f, 0.f);
    if (slots > 0) { j = __ldcs(J + e0); wp = __ldcs(WP + e0); if (ALPHA) wb = __ldcs(WB + e0); }
#pragma unroll 1
    for (int s = 0; s < slots; ++s) {
        const float4* a4 = reinterpret_cast<const float4*>(A_s + (size_t)j * NF);
        float2 Aj[NF / 2];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float4 t = __ldg(a4 + k); Aj[2*k] = make_float2(t.x, t.y); Aj[2*k+1] = make_float2(t.z, t.w); }
        const float2 wpc = wp, wbc = wb;
        if (s + 1 < slots) { const int64_t e = e0 + (int64_t)(s + 1) * 32; j = __ldcs(J + e); wp = __ldcs(WP + e); if (ALPHA) wb = __ldcs(WB + e); }
        if (ALPHA) {
            const float alive = Aj[1].y > 0.1f ? 1.f : 0.f;                         // channel 3
            blur = fmaf(alive * wbc.x, wbc.y, blur);
        }
        const float2 wx = make_float2(wpc.x, wpc.x), wy = make_float2(wpc.y, wpc.y);
#pragma unroll
        for (int f = 0; f < NF / 2; ++f) {
            const float2 dA = fsub2(Aj[f], Ai[f]);
            gx[f] = ffma2(dA, wx, gx[f]);
            gy[f] = ffma2(dA, wy, gy[f]);
        }
    }
    if (p >= n) return;
    float4* y4 = reinterpret_cast<float4*>(Y_s + (size_t)p * 2 * NF);
#pragma unroll
    for (int q = 0; q < 4; ++q) y4[q] = make_float4(gx[2*q].x * c.kg, gx[2*q].y * c.kg, gx[2*q+1].x * c.kg, gx[2*q+1].y * c.kg);
#pragma unroll
    for (int q = 0; q < 4; ++q) y4[4 + q] = make_float4(gy[2*q].x * c.kg, gy[2*q].y * c.kg, gy[2*q+1].x * c.kg, gy[2*q+1].y * c.kg);
    flags_s[p] = ALPHA ? (blur * c.cw > 0.1f ? 1 : 0) : FLG[p];
}

// Gather adjoint of the perception sweep from the pair lists (same output as step_adjoint_kernel).
// FOUR lanes per particle, lane part c owns channels 4c..4c+3: the 128-byte neighbour row is fetched by 4 adjacent lanes
// as two contiguous 64-byte pieces and a warp-wide load touches 8 cache lines instead of 32.  With one lane per particle
// this sweep was bound by the L1 tag stage (one line lookup per cycle per SM): 0.52 ms, slower than the cell sweep;
// this form runs at 87 % of the L1 data-pipe throughput (ncu).  A CTA of 128 threads = one list block of 32 particles.
// Per channel the accumulation order is the slot order, so the result stays bit-identical to the gather kernel.
__global__ void __launch_bounds__(128)
step_adjoint_pairs_kernel(int64_t n, const int32_t* __restrict__ warp_slots, const int64_t* __restrict__ warp_base,
                          const int32_t* __restrict__ J, const float2* __restrict__ WA, const float2* __restrict__ SXY,
                          const float4* __restrict__ xv4, const float* __restrict__ dY_s, const float* __restrict__ dAdir_s,
                          StepConsts c, float* __restrict__ dA_s) {
    const int64_t w = blockIdx.x;
    const int part = threadIdx.x & 3;
    const int pl = threadIdx.x >> 2;
    const int64_t p = w * 32 + pl;
    const int slots = warp_slots[w];
    const int64_t e0 = warp_base[w] + pl;
    const float4* __restrict__ G4 = reinterpret_cast<const float4*>(dY_s);      // row = 8 float4: x part 0..3, y part 4..7
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int PB = 4;                                 // slots per batch: list loads, then gathers, then arithmetic
    auto accumulate = [&](const float4 a, const float4 b, const float2 wa) {
        const float2 wx = make_float2(wa.x, wa.x), wy = make_float2(wa.y, wa.y);
        float2 t;
        t = ffma2(wx, make_float2(a.x, a.y), ffma2(wy, make_float2(b.x, b.y), make_float2(acc.x, acc.y))); acc.x = t.x; acc.y = t.y;
        t = ffma2(wx, make_float2(a.z, a.w), ffma2(wy, make_float2(b.z, b.w), make_float2(acc.z, acc.w))); acc.z = t.x; acc.w = t.y;
    };
    int s = 0;
#pragma unroll 1
    for (; s + PB <= slots; s += PB) {
        int j[PB]; float2 wa[PB]; float4 a[PB], b[PB];
#pragma unroll
        for (int u = 0; u < PB; ++u) { const int64_t e = e0 + (int64_t)(s + u) * 32; j[u] = __ldcs(J + e); wa[u] = __ldcs(WA + e); }
#pragma unroll
        for (int u = 0; u < PB; ++u) { a[u] = __ldg(G4 + (size_t)j[u] * 8 + part); b[u] = __ldg(G4 + (size_t)j[u] * 8 + 4 + part); }
#pragma unroll
        for (int u = 0; u < PB; ++u) accumulate(a[u], b[u], wa[u]);
    }
    for (; s < slots; ++s) {
        const int64_t e = e0 + (int64_t)s * 32;
        const int j = __ldcs(J + e);
        accumulate(__ldg(G4 + (size_t)j * 8 + part), __ldg(G4 + (size_t)j * 8 + 4 + part), __ldcs(WA + e));
    }
    if (p >= n) return;
    const float vi = xv4[p].w;
    const float2 sxy = SXY[p];
    const float4 Gx = __ldg(G4 + (size_t)p * 8 + part), Gy = __ldg(G4 + (size_t)p * 8 + 4 + part);
    const float4 dir = __ldg(reinterpret_cast<const float4*>(dAdir_s) + (size_t)p * 4 + part);
    float4 r;
    r.x = dir.x - c.kg * fmaf(vi, acc.x, fmaf(Gx.x, sxy.x, Gy.x * sxy.y));
    r.y = dir.y - c.kg * fmaf(vi, acc.y, fmaf(Gx.y, sxy.x, Gy.y * sxy.y));
    r.z = dir.z - c.kg * fmaf(vi, acc.z, fmaf(Gx.z, sxy.x, Gy.z * sxy.y));
    r.w = dir.w - c.kg * fmaf(vi, acc.w, fmaf(Gx.w, sxy.x, Gy.w * sxy.y));
    reinterpret_cast<float4*>(dA_s)[(size_t)p * 4 + part] = r;
}

}  // namespace sphnca

extern "C" {
using namespace sphnca;

int32_t sphnca_pairs_count(const sphnca_grid* g, const float* xv4, int32_t* warp_slots, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv); if (rc) return rc; }
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(xv4 && warp_slots, "NULL argument");
    const int blocks = (int)((gv.n + 127) / 128);
    const float h2 = (float)(g->h * g->h);
    if (gv.dim == 3) SPHNCA_LAUNCH(pairs_count_kernel<3>, blocks, 128, 0, as_stream(stream_), gv, (const float4*)xv4, h2, warp_slots);
    else SPHNCA_LAUNCH(pairs_count_kernel<2>, blocks, 128, 0, as_stream(stream_), gv, (const float4*)xv4, h2, warp_slots);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_pairs_fill(const sphnca_grid* g, const float* xv4, const int32_t* warp_slots, const int64_t* warp_base,
                          int32_t* J, float* WP, float* WB, float* WA, float* SXY, uint8_t* FLG, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv); if (rc) return rc; }
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(xv4 && warp_slots && warp_base && J && WP && WB && WA && SXY && FLG, "NULL argument");
    const float cw = (float)poly6_norm(g->dim, g->h);
    const int blocks = (int)((gv.n + 127) / 128);
    const float h = (float)g->h, h2 = (float)(g->h * g->h);
    if (gv.dim == 3) SPHNCA_LAUNCH(pairs_fill_kernel<3>, blocks, 128, 0, as_stream(stream_), gv, (const float4*)xv4, h, h2, warp_slots, warp_base,
                                   J, (float2*)WP, (float2*)WB, (float2*)WA, (float2*)SXY, cw, FLG);
    else SPHNCA_LAUNCH(pairs_fill_kernel<2>, blocks, 128, 0, as_stream(stream_), gv, (const float4*)xv4, h, h2, warp_slots, warp_base,
                       J, (float2*)WP, (float2*)WB, (float2*)WA, (float2*)SXY, cw, FLG);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_step_perceive_pairs(const sphnca_grid* g, const int32_t* warp_slots, const int64_t* warp_base, const int32_t* J,
                                   const float* WP, const float* WB, const uint8_t* FLG, const float* A_s,
                                   const sphnca_step_params* p, float* Y_s, uint8_t* flags_s, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv); if (rc) return rc; }
    StepConsts c; { int rc = make_consts(g, p, &c); if (rc) return rc; }
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(warp_slots && warp_base && J && WP && WB && FLG && A_s && Y_s && flags_s, "NULL argument");
    const int blocks = (int)((gv.n + 127) / 128);
    if (c.use_alpha) SPHNCA_LAUNCH(step_perceive_pairs_kernel<true>, blocks, 128, 0, as_stream(stream_), gv.n, warp_slots, warp_base, J,
                                   (const float2*)WP, (const float2*)WB, FLG, A_s, c, Y_s, flags_s);
    else SPHNCA_LAUNCH(step_perceive_pairs_kernel<false>, blocks, 128, 0, as_stream(stream_), gv.n, warp_slots, warp_base, J,
                       (const float2*)WP, (const float2*)WB, FLG, A_s, c, Y_s, flags_s);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_step_adjoint_pairs(const sphnca_grid* g, const int32_t* warp_slots, const int64_t* warp_base, const int32_t* J,
                                  const float* WA, const float* SXY, const float* xv4, const float* dY_s, const float* dAdir_s,
                                  const sphnca_step_params* p, float* dA_s, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv); if (rc) return rc; }
    StepConsts c; { int rc = make_consts(g, p, &c); if (rc) return rc; }
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(warp_slots && warp_base && J && WA && SXY && xv4 && dY_s && dAdir_s && dA_s, "NULL argument");
    const int blocks = (int)((gv.n + 31) / 32);           // one CTA per list block of 32 particles, 4 lanes per particle
    SPHNCA_LAUNCH(step_adjoint_pairs_kernel, blocks, 128, 0, as_stream(stream_), gv.n, warp_slots, warp_base, J, (const float2*)WA,
                  (const float2*)SXY, (const float4*)xv4, dY_s, dAdir_s, c, dA_s);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

}  // extern "C"
