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
//   FLG uint8   per particle: life mask of an all-alive field (what the perception reports without an alpha channel)
//
// Layout: per GROUP of 8 consecutive cell-sorted particles, [slot][8]; a group's slot count is the largest neighbour
// count of its particles rounded up to a multiple of 4, shorter lists are padded with zero-weight self pairs.
// Sweeps: one warp per group, FOUR lanes per particle (lane part c owns channels 4c..4c+3).
//   list entries   one warp-wide load per array fetches 4 slots x 8 particles = 32 consecutive entries (128/256 bytes,
//                  fully coalesced, streaming); lane l holds the entry (slot l/8, particle l%8) and hands it to the 4
//                  lanes of that particle by shuffle;
//   row gathers    the 4 lanes of a particle fetch a neighbour row as one contiguous 64-byte piece, so a warp-wide load
//                  touches 8 cache lines instead of 32.
// Why: with one lane per particle the sweeps are bound by the L1 tag stage (one line lookup per cycle per SM; 0.52 ms for
// the adjoint, slower than the cell sweep), and with per-particle list loads in the 4-lane mapping each warp reads
// 32-byte crumbs of the lists.  The loops are warp-uniform: no divergence, no in-radius tests, no geometry.
// Entries are appended in the candidate order of for_each_candidate and hold the very floats the gather kernels compute;
// per channel the accumulation order is the slot order, so the sums are bit-identical to step_perceive_kernel /
// step_adjoint_kernel.  Cost: 28 B per pair (about 1 KB per particle), one count + one fill sweep per grid.
#include "step_common.cuh"

namespace sphnca {

constexpr int PG = 8;                                     // particles per list group

template <int D>
__global__ void __launch_bounds__(128)
pairs_count_kernel(GridView g, const float4* __restrict__ xv4, float h2, int32_t* __restrict__ group_slots) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    if (p < g.n) {
        const float4 xi = xv4[p];
        for_each_candidate<D>(g, g.sorted_cell[p], [&](int q) {
            float rx, ry;
            float4 xj = __ldg(&xv4[q]);
            if (g.periodic) xj = nearest_image(g, xi, xj);
            if (pair_d2<D>(xi, xj, rx, ry) < h2) ++cnt;
        });
    }
    // max over the 8 lanes of the group
    cnt = max(cnt, __shfl_xor_sync(0xffffffffu, cnt, 1));
    cnt = max(cnt, __shfl_xor_sync(0xffffffffu, cnt, 2));
    cnt = max(cnt, __shfl_xor_sync(0xffffffffu, cnt, 4));
    if ((threadIdx.x & (PG - 1)) == 0 && p < g.n) group_slots[p / PG] = (cnt + 3) & ~3;
}

template <int D>
__global__ void __launch_bounds__(128)
pairs_fill_kernel(GridView g, const float4* __restrict__ xv4, float h, float h2, const int32_t* __restrict__ group_slots,
                  const int64_t* __restrict__ group_base, int32_t* __restrict__ J, float2* __restrict__ WP,
                  float2* __restrict__ WB, float2* __restrict__ WA, float2* __restrict__ SXY, float cw,
                  uint8_t* __restrict__ FLG, float* __restrict__ WPZ) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t grp = p / PG;
    if (grp * PG >= g.n) return;
    const int slots = group_slots[grp];
    int64_t e = group_base[grp] + (p & (PG - 1));
    int cnt = 0;
    const int64_t self = p < g.n ? p : g.n - 1;
    if (p < g.n) {
        const float4 xi = xv4[p];
        float sx = 0.f, sy = 0.f, blur1 = 0.f;
        for_each_candidate<D>(g, g.sorted_cell[p], [&](int q) {
            float rx, ry;
            float4 xj = __ldg(&xv4[q]);
            float self_ratio = 1.f;
            if (g.periodic) {
                const float4 xo = xj;
                xj = nearest_image(g, xi, xj);                   // the ghost copy's position (periodic planes, nca.py:200-227)
                if (g.image_volume != nullptr) {
                    int tx, ty;
                    image_shift(xo, xj, tx, ty);
                    if (tx | ty) {
                        // a neighbour through the wrap is the reference's GHOST copy (tx, ty) of q, which carries the volume of its own
                        // truncated neighbourhood; its counterpart in q's list is the copy (-tx, -ty) of this particle, whose volume
                        // multiplies this row's share of the adjoint (that ghost row's gradient is gathered back into this row)
                        xj.w = __ldg(g.image_volume + (size_t)q * 9 + (tx + 1) * 3 + (ty + 1));
                        self_ratio = __ldg(g.image_volume + (size_t)p * 9 + (1 - tx) * 3 + (1 - ty)) / xi.w;
                    }
                }
            }
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
                    if (self_ratio != 1.f) wa = make_float2(wa.x * self_ratio, wa.y * self_ratio);
                }
                J[e] = q; WP[e] = wp; WA[e] = wa;
                if (WB != nullptr) WB[e] = make_float2(xj.w, t * t * t);
                if (WPZ != nullptr) {                                    // z component of the perception weight (tangent-frame sweeps)
                    float wz = 0.f;
                    if (D == 3 && d2 > 0.f) { const float inv = rsqrtf(d2); const float hd = h - d2 * inv; wz = 3.f * hd * hd * inv * xj.w * (xj.z - xi.z); }
                    WPZ[e] = wz;
                }
                e += PG; ++cnt;
            }
        });
        SXY[p] = make_float2(sx, sy);
        FLG[p] = blur1 * cw > 0.1f ? 1 : 0;
    }
    for (; cnt < slots; ++cnt, e += PG) {                 // zero-weight self pairs up to the group's slot count
        J[e] = (int32_t)self; WP[e] = make_float2(0.f, 0.f); WA[e] = make_float2(0.f, 0.f);
        if (WB != nullptr) WB[e] = make_float2(0.f, 0.f);
        if (WPZ != nullptr) WPZ[e] = 0.f;
    }
}

// Perception sweep + previous life mask from the pair lists (same outputs as step_perceive_kernel).
// ALPHA = false (no alpha channel): the life mask does not depend on the state; it was evaluated once by the fill kernel
// (FLG) and the blur weights WB are not read at all.
template <bool ALPHA>
__global__ void __launch_bounds__(128)
step_perceive_pairs_kernel(int64_t n, const int32_t* __restrict__ group_slots, const int64_t* __restrict__ group_base,
                           const int32_t* __restrict__ J, const float2* __restrict__ WP, const float2* __restrict__ WB,
                           const uint8_t* __restrict__ FLG, const float* __restrict__ A_s, StepConsts c,
                           float* __restrict__ Y_s, uint8_t* __restrict__ flags_s) {
    const int64_t grp = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (grp * PG >= n) return;
    const int lane = threadIdx.x & 31, part = lane & 3, q = lane >> 2;
    const int64_t p = grp * PG + q;
    const int64_t pi = p < n ? p : n - 1;
    const int batches = group_slots[grp] >> 2;            // 4 slots per batch
    const int64_t e0 = group_base[grp] + lane;
    const float4* __restrict__ A4 = reinterpret_cast<const float4*>(A_s);
    const float4 Ai = __ldg(A4 + (size_t)pi * 4 + part);
    float4 gx = make_float4(0.f, 0.f, 0.f, 0.f), gy = make_float4(0.f, 0.f, 0.f, 0.f);
    float blur = 0.f;
    // the entries of the next batch are in flight during the arithmetic of the current one
    int jv = 0; float2 wp = make_float2(0.f, 0.f), wb = make_float2(0.f, 0.f);
    if (batches > 0) { jv = __ldcs(J + e0); wp = __ldcs(WP + e0); if (ALPHA) wb = __ldcs(WB + e0); }
#pragma unroll 1
    for (int b = 0; b < batches; ++b) {
        float4 Aj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) Aj[u] = __ldg(A4 + (size_t)__shfl_sync(0xffffffffu, jv, u * PG + q) * 4 + part);
        const float2 wpc = wp, wbc = wb;
        if (b + 1 < batches) { const int64_t e = e0 + (int64_t)(b + 1) * 32; jv = __ldcs(J + e); wp = __ldcs(WP + e); if (ALPHA) wb = __ldcs(WB + e); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int src = u * PG + q;
            const float wxs = __shfl_sync(0xffffffffu, wpc.x, src), wys = __shfl_sync(0xffffffffu, wpc.y, src);
            if (ALPHA) {
                const float bv = __shfl_sync(0xffffffffu, wbc.x, src), bt = __shfl_sync(0xffffffffu, wbc.y, src);
                const float alive = Aj[u].w > 0.1f ? 1.f : 0.f;                    // channel 3: meaningful in part 0 only
                blur = fmaf(alive * bv, bt, blur);
            }
            const float2 wx = make_float2(wxs, wxs), wy = make_float2(wys, wys);
            const float2 d0 = fsub2(make_float2(Aj[u].x, Aj[u].y), make_float2(Ai.x, Ai.y));
            const float2 d1 = fsub2(make_float2(Aj[u].z, Aj[u].w), make_float2(Ai.z, Ai.w));
            float2 t;
            t = ffma2(d0, wx, make_float2(gx.x, gx.y)); gx.x = t.x; gx.y = t.y;
            t = ffma2(d1, wx, make_float2(gx.z, gx.w)); gx.z = t.x; gx.w = t.y;
            t = ffma2(d0, wy, make_float2(gy.x, gy.y)); gy.x = t.x; gy.y = t.y;
            t = ffma2(d1, wy, make_float2(gy.z, gy.w)); gy.z = t.x; gy.w = t.y;
        }
    }
    if (p >= n) return;
    float4* y4 = reinterpret_cast<float4*>(Y_s + (size_t)p * 2 * NF);
    y4[part] = make_float4(gx.x * c.kg, gx.y * c.kg, gx.z * c.kg, gx.w * c.kg);
    y4[4 + part] = make_float4(gy.x * c.kg, gy.y * c.kg, gy.z * c.kg, gy.w * c.kg);
    if (part == 0) flags_s[p] = ALPHA ? (blur * c.cw > 0.1f ? 1 : 0) : FLG[p];
}

// Life mask from the pair lists (models with an alpha channel): blur_i = cw * sum_j alive_j v_j (h^2 - d_ij^2)^3 > 0.1 over the
// list (J, WB = {v_j, (h^2 - d^2)^3}), one thread per particle -- the F = 1 blur of reference nca.py:71-79 without its 16-channel
// company: 12 list bytes + one byte gather per neighbour.  Same terms in the same order as the blur inside
// step_perceive_pairs_kernel<true> and as step_mask_kernel: bit-identical masks.
//   MODE 0  previous mask of a step: alive_in = alive bytes of the CURRENT state; flags_s = bit 0.
//   MODE 1  new mask of a step: alive_in = alive bytes of the unmasked NEW state; flags_s |= bit 1; rows that are not living
//           (prev & new) are multiplied by 0; alive_out = alive bytes of the masked new state (the next step's MODE 0 input).
template <int MODE>
__global__ void __launch_bounds__(256)
step_mask_pairs_kernel(int64_t n, const int32_t* __restrict__ group_slots, const int64_t* __restrict__ group_base,
                       const int32_t* __restrict__ J, const float2* __restrict__ WB, const uint8_t* __restrict__ alive_in, float cw,
                       uint8_t* __restrict__ flags_s, float* __restrict__ An_s, uint8_t* __restrict__ alive_out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int64_t grp = p >> 3;
    const int slots = group_slots[grp];
    int64_t e = group_base[grp] + (p & 7);
    float blur = 0.f;
#pragma unroll 4
    for (int s = 0; s < slots; ++s, e += PG) {
        const int j = __ldcs(J + e);
        const float2 wb = __ldcs(WB + e);
        const float alive = __ldg(alive_in + j) ? 1.f : 0.f;
        blur = fmaf(alive * wb.x, wb.y, blur);
    }
    const bool on = blur * cw > 0.1f;
    if (MODE == 0) {
        flags_s[p] = on ? 1 : 0;
    } else {
        uint8_t fl = flags_s[p];
        if (on) fl |= 2;
        flags_s[p] = fl;
        const bool living = (fl & 1) && on;
        if (!living) {
            float4* a4 = reinterpret_cast<float4*>(An_s + (size_t)p * NF);
#pragma unroll
            for (int k = 0; k < 4; ++k) { float4 t = a4[k]; a4[k] = make_float4(t.x * 0.f, t.y * 0.f, t.z * 0.f, t.w * 0.f); }
        }
        if (alive_out != nullptr) alive_out[p] = living ? alive_in[p] : (uint8_t)0;
    }
}

// Perception in a per-particle TANGENT FRAME (surface point clouds, reference nca.py:325-336,364: `feature_process_tangent` projects the
// 3-D gradient of every channel onto (T_i, B_i = N_i x T_i) after the sweep): the projection is linear, so it is applied to the pair
// WEIGHT instead -- w_t = w . T_i, w_b = w . B_i, 6 FMAs per pair in the lane that loaded the entry -- and the sweep itself is the
// planar one: Y = [sum_j (A_j - A_i) w_t | sum_j (A_j - A_i) w_b].  frame_s [N][8] = {T, 0, B, 0} in cell-sorted order.
template <bool ALPHA>
__global__ void __launch_bounds__(128)
step_perceive_pairs_frame_kernel(int64_t n, const int32_t* __restrict__ group_slots, const int64_t* __restrict__ group_base,
                                 const int32_t* __restrict__ J, const float2* __restrict__ WP, const float* __restrict__ WPZ,
                                 const float2* __restrict__ WB, const uint8_t* __restrict__ FLG, const float4* __restrict__ frame_s,
                                 const float* __restrict__ A_s, StepConsts c, float* __restrict__ Y_s, uint8_t* __restrict__ flags_s) {
    const int64_t grp = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (grp * PG >= n) return;
    const int lane = threadIdx.x & 31, part = lane & 3, q = lane >> 2;
    const int64_t p = grp * PG + q;
    const int64_t pi = p < n ? p : n - 1;
    const int batches = group_slots[grp] >> 2;            // 4 slots per batch
    const int64_t e0 = group_base[grp] + lane;
    // the frame of the particle whose list entries THIS lane loads (column lane % 8 of the group)
    const int64_t pe = grp * PG + (lane & 7) < n ? grp * PG + (lane & 7) : n - 1;
    const float4 T = __ldg(frame_s + pe * 2), B = __ldg(frame_s + pe * 2 + 1);
    const float4* __restrict__ A4 = reinterpret_cast<const float4*>(A_s);
    const float4 Ai = __ldg(A4 + (size_t)pi * 4 + part);
    float4 gx = make_float4(0.f, 0.f, 0.f, 0.f), gy = make_float4(0.f, 0.f, 0.f, 0.f);
    float blur = 0.f;
    int jv = 0; float2 wp = make_float2(0.f, 0.f), wb = make_float2(0.f, 0.f); float wz = 0.f;
    if (batches > 0) { jv = __ldcs(J + e0); wp = __ldcs(WP + e0); wz = __ldcs(WPZ + e0); if (ALPHA) wb = __ldcs(WB + e0); }
#pragma unroll 1
    for (int b = 0; b < batches; ++b) {
        float4 Aj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) Aj[u] = __ldg(A4 + (size_t)__shfl_sync(0xffffffffu, jv, u * PG + q) * 4 + part);
        // projected weights of this lane's entry
        const float wt = fmaf(wz, T.z, fmaf(wp.y, T.y, wp.x * T.x)), wbn = fmaf(wz, B.z, fmaf(wp.y, B.y, wp.x * B.x));
        const float2 wbc = wb;
        if (b + 1 < batches) { const int64_t e = e0 + (int64_t)(b + 1) * 32; jv = __ldcs(J + e); wp = __ldcs(WP + e); wz = __ldcs(WPZ + e); if (ALPHA) wb = __ldcs(WB + e); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int src = u * PG + q;
            const float wxs = __shfl_sync(0xffffffffu, wt, src), wys = __shfl_sync(0xffffffffu, wbn, src);
            if (ALPHA) {
                const float bv = __shfl_sync(0xffffffffu, wbc.x, src), bt = __shfl_sync(0xffffffffu, wbc.y, src);
                const float alive = Aj[u].w > 0.1f ? 1.f : 0.f;                    // channel 3: meaningful in part 0 only
                blur = fmaf(alive * bv, bt, blur);
            }
            const float2 wx = make_float2(wxs, wxs), wy = make_float2(wys, wys);
            const float2 d0 = fsub2(make_float2(Aj[u].x, Aj[u].y), make_float2(Ai.x, Ai.y));
            const float2 d1 = fsub2(make_float2(Aj[u].z, Aj[u].w), make_float2(Ai.z, Ai.w));
            float2 t;
            t = ffma2(d0, wx, make_float2(gx.x, gx.y)); gx.x = t.x; gx.y = t.y;
            t = ffma2(d1, wx, make_float2(gx.z, gx.w)); gx.z = t.x; gx.w = t.y;
            t = ffma2(d0, wy, make_float2(gy.x, gy.y)); gy.x = t.x; gy.y = t.y;
            t = ffma2(d1, wy, make_float2(gy.z, gy.w)); gy.z = t.x; gy.w = t.y;
        }
    }
    if (p >= n) return;
    float4* y4 = reinterpret_cast<float4*>(Y_s + (size_t)p * 2 * NF);
    y4[part] = make_float4(gx.x * c.kg, gx.y * c.kg, gx.z * c.kg, gx.w * c.kg);
    y4[4 + part] = make_float4(gy.x * c.kg, gy.y * c.kg, gy.z * c.kg, gy.w * c.kg);
    if (part == 0) flags_s[p] = ALPHA ? (blur * c.cw > 0.1f ? 1 : 0) : FLG[p];
}

// Perception for the FIRED particles only (fire-list rollouts without an alpha channel, firelist.cu): the perception vector
// feeds nothing but the MLP, and the MLP runs over the fired rows only, so the sweep of a not-fired particle is dead work.
// One warp takes 8 consecutive list positions (4 lanes per particle as above); the 4 lanes of a particle fetch 4 slots of its
// own list column (particles of one group share the 32-byte sectors of a slot), entries travel inside the quad by shuffle.
// Same slot order and arithmetic as the full sweep: Y_c[j] is bit-identical to Y_s[rows[j]].  Also writes the compact copy of
// the state row, so the MLP kernel reads [A_c | Y_c] rows contiguously by list position.
__global__ void __launch_bounds__(128)
step_perceive_pairs_fired_kernel(const int32_t* __restrict__ rows, const int32_t* __restrict__ count_ptr,
                                 const int32_t* __restrict__ group_slots, const int64_t* __restrict__ group_base,
                                 const int32_t* __restrict__ J, const float2* __restrict__ WP, const float* __restrict__ A_s,
                                 StepConsts c, float* __restrict__ Y_c, float* __restrict__ A_c) {
    const int cnt = __ldg(count_ptr);
    const int64_t j0 = ((int64_t)blockIdx.x * 4 + (threadIdx.x >> 5)) * PG;
    if (j0 >= cnt) return;
    const int lane = threadIdx.x & 31, part = lane & 3, q = lane >> 2, quad0 = lane & ~3;
    const int64_t j = j0 + q;
    const bool valid = j < cnt;
    const int64_t p = __ldg(rows + (valid ? j : (int64_t)cnt - 1));
    const int64_t grp = p >> 3;
    const int nb = group_slots[grp] >> 2;                 // batches of 4 slots of this particle's group
    int maxb = nb;
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) maxb = max(maxb, __shfl_xor_sync(0xffffffffu, maxb, o));
    const int64_t e0 = group_base[grp] + (p & 7) + (int64_t)part * PG;      // slot `part` of batch 0, this particle's column
    const float4* __restrict__ A4 = reinterpret_cast<const float4*>(A_s);
    const float4 Ai = __ldg(A4 + (size_t)p * 4 + part);
    float4 gx = make_float4(0.f, 0.f, 0.f, 0.f), gy = make_float4(0.f, 0.f, 0.f, 0.f);
    // past the end of a shorter list: zero-weight self pairs (adding 0 * (A_i - A_i) leaves the sums bit-identical)
    int jv = (int)p; float2 wp = make_float2(0.f, 0.f);
    if (nb > 0) { jv = __ldcs(J + e0); wp = __ldcs(WP + e0); }
#pragma unroll 1
    for (int b = 0; b < maxb; ++b) {
        float4 Aj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) Aj[u] = __ldg(A4 + (size_t)__shfl_sync(0xffffffffu, jv, quad0 + u) * 4 + part);
        const float2 wpc = wp;
        if (b + 1 < nb) { const int64_t e = e0 + (int64_t)(b + 1) * 32; jv = __ldcs(J + e); wp = __ldcs(WP + e); }
        else { jv = (int)p; wp = make_float2(0.f, 0.f); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float wxs = __shfl_sync(0xffffffffu, wpc.x, quad0 + u), wys = __shfl_sync(0xffffffffu, wpc.y, quad0 + u);
            const float2 wx = make_float2(wxs, wxs), wy = make_float2(wys, wys);
            const float2 d0 = fsub2(make_float2(Aj[u].x, Aj[u].y), make_float2(Ai.x, Ai.y));
            const float2 d1 = fsub2(make_float2(Aj[u].z, Aj[u].w), make_float2(Ai.z, Ai.w));
            float2 t;
            t = ffma2(d0, wx, make_float2(gx.x, gx.y)); gx.x = t.x; gx.y = t.y;
            t = ffma2(d1, wx, make_float2(gx.z, gx.w)); gx.z = t.x; gx.w = t.y;
            t = ffma2(d0, wy, make_float2(gy.x, gy.y)); gy.x = t.x; gy.y = t.y;
            t = ffma2(d1, wy, make_float2(gy.z, gy.w)); gy.z = t.x; gy.w = t.y;
        }
    }
    if (!valid) return;
    float4* y4 = reinterpret_cast<float4*>(Y_c + (size_t)j * 2 * NF);
    y4[part] = make_float4(gx.x * c.kg, gx.y * c.kg, gx.z * c.kg, gx.w * c.kg);
    y4[4 + part] = make_float4(gy.x * c.kg, gy.y * c.kg, gy.z * c.kg, gy.w * c.kg);
    reinterpret_cast<float4*>(A_c + (size_t)j * NF)[part] = Ai;
}

__device__ float4 g_zero_row[8];
// 32 bytes in one load: the x and y quarters of an interleaved dL/dY row (sphnca_step_params.dy_interleaved)
__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}      // 128 bytes of zeros: the dL/dY row of every particle that did not fire

// Gather adjoint of the perception sweep from the pair lists (same output as step_adjoint_kernel).
// FIRED (fire-list rollouts, firelist.cu): dL/dY of a particle that did not fire is exactly zero and its direct dL/dA part is
// (living ? dL/dA_new : 0), so both are taken from the flag byte instead of from memory: the 128-byte row gather of a not-fired
// neighbour (half of them at fire rate 0.5 -- this kernel is bound by the L1 data pipe) is skipped, and nobody has to write
// those rows of dY_s / dAdir_s first.
// ILV: the dL/dY rows are [x quarter | y quarter] pairs (sphnca_step_params.dy_interleaved) -- a template parameter, not a branch
// in the loop: with the branch the four gathers of a batch sit in four basic blocks and are no longer issued back to back.
template <bool FIRED, bool ILV>
__global__ void __launch_bounds__(128)
step_adjoint_pairs_kernel(int64_t n, const int32_t* __restrict__ group_slots, const int64_t* __restrict__ group_base,
                          const int32_t* __restrict__ J, const float2* __restrict__ WA, const float2* __restrict__ SXY,
                          const float4* __restrict__ xv4, const float* __restrict__ dY_s, const float* __restrict__ dAdir_s,
                          StepConsts c, float* __restrict__ dA_s, const uint8_t* __restrict__ flags_s,
                          const float* __restrict__ dAn_s) {
    const int64_t grp = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (grp * PG >= n) return;
    const int lane = threadIdx.x & 31, part = lane & 3, q = lane >> 2;
    const int64_t p = grp * PG + q;
    const int batches = group_slots[grp] >> 2;
    const float4* __restrict__ G4 = reinterpret_cast<const float4*>(dY_s);      // row = 8 float4: x part 0..3, y part 4..7 (or x|y pairs)
    constexpr bool ilv = ILV;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // Every lane of a quad reads its particle's list entries itself (same address: a broadcast inside the load; a warp = one group,
    // so a slot's 8 entries are 32 / 64 contiguous bytes) instead of one lane per entry + 3 shuffles per pair: shuffles share the
    // LSU pipe this sweep is bound by (0.1856 -> 0.1783 ms).  Not so in the fired-rows perception sweep, whose quads belong to
    // different groups: there the direct reads would touch 4x the lines.
    const int64_t eq = group_base[grp] + q;
    int jn[4]; float2 wn[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { jn[u] = 0; wn[u] = make_float2(0.f, 0.f); }
    if (batches > 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { jn[u] = __ldcs(J + eq + u * PG); wn[u] = __ldcs(WA + eq + u * PG); }
    }
#pragma unroll 1
    for (int b = 0; b < batches; ++b) {
        float4 a[4], bb[4];
        float2 wc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = jn[u];
            wc[u] = wn[u];
            const size_t row = (size_t)j * 8;
            const float4* src = (FIRED && !(__ldg(flags_s + j) & 4)) ? g_zero_row : G4 + row;
            if (ilv) ldg256(src + 2 * part, a[u], bb[u]);     // one line lookup per row instead of two (0.174 -> 0.162 ms)
            else { a[u] = __ldg(src + part); bb[u] = __ldg(src + 4 + part); }
        }
        if (b + 1 < batches) {
            const int64_t e = eq + (int64_t)(b + 1) * 32;
#pragma unroll
            for (int u = 0; u < 4; ++u) { jn[u] = __ldcs(J + e + u * PG); wn[u] = __ldcs(WA + e + u * PG); }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float2 wx = make_float2(wc[u].x, wc[u].x), wy = make_float2(wc[u].y, wc[u].y);
            float2 t;
            t = ffma2(wx, make_float2(a[u].x, a[u].y), ffma2(wy, make_float2(bb[u].x, bb[u].y), make_float2(acc.x, acc.y))); acc.x = t.x; acc.y = t.y;
            t = ffma2(wx, make_float2(a[u].z, a[u].w), ffma2(wy, make_float2(bb[u].z, bb[u].w), make_float2(acc.z, acc.w))); acc.z = t.x; acc.w = t.y;
        }
    }
    if (p >= n) return;
    const float vi = xv4[p].w;
    const float2 sxy = SXY[p];
    float4 Gx, Gy, dir;
    const uint8_t fl_ = FIRED ? flags_s[p] : (uint8_t)4;
    if (fl_ & 4) {
        if (ilv) ldg256(G4 + (size_t)p * 8 + 2 * part, Gx, Gy);
        else { Gx = __ldg(G4 + (size_t)p * 8 + part); Gy = __ldg(G4 + (size_t)p * 8 + 4 + part); }
        dir = __ldg(reinterpret_cast<const float4*>(dAdir_s) + (size_t)p * 4 + part);
    } else {
        Gx = make_float4(0.f, 0.f, 0.f, 0.f); Gy = Gx;
        dir = ((fl_ & 1) && (fl_ & 2)) ? __ldg(reinterpret_cast<const float4*>(dAn_s) + (size_t)p * 4 + part) : Gx;
    }
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

int32_t sphnca_pairs_count(const sphnca_grid* g, const float* xv4, int32_t* group_slots, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv, true); if (rc) return rc; }
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(xv4 && group_slots, "NULL argument");
    const int blocks = (int)((gv.n + 127) / 128);
    const float h2 = (float)(g->h * g->h);
    if (gv.dim == 3) SPHNCA_LAUNCH(pairs_count_kernel<3>, blocks, 128, 0, as_stream(stream_), gv, (const float4*)xv4, h2, group_slots);
    else SPHNCA_LAUNCH(pairs_count_kernel<2>, blocks, 128, 0, as_stream(stream_), gv, (const float4*)xv4, h2, group_slots);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_pairs_fill(const sphnca_grid* g, const float* xv4, const int32_t* group_slots, const int64_t* group_base,
                          int32_t* J, float* WP, float* WB, float* WA, float* SXY, uint8_t* FLG, float* WPZ, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv, true); if (rc) return rc; }
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(xv4 && group_slots && group_base && J && WP && WA && SXY && FLG, "NULL argument");   // WB may be NULL
    const float cw = (float)poly6_norm(g->dim, g->h);
    const int blocks = (int)((gv.n + 127) / 128);
    const float h = (float)g->h, h2 = (float)(g->h * g->h);
    if (gv.dim == 3) SPHNCA_LAUNCH(pairs_fill_kernel<3>, blocks, 128, 0, as_stream(stream_), gv, (const float4*)xv4, h, h2, group_slots, group_base,
                                   J, (float2*)WP, (float2*)WB, (float2*)WA, (float2*)SXY, cw, FLG, WPZ);
    else SPHNCA_LAUNCH(pairs_fill_kernel<2>, blocks, 128, 0, as_stream(stream_), gv, (const float4*)xv4, h, h2, group_slots, group_base,
                       J, (float2*)WP, (float2*)WB, (float2*)WA, (float2*)SXY, cw, FLG, WPZ);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_step_perceive_pairs(const sphnca_grid* g, const int32_t* group_slots, const int64_t* group_base, const int32_t* J,
                                   const float* WP, const float* WB, const uint8_t* FLG, const float* A_s,
                                   const sphnca_step_params* p, float* Y_s, uint8_t* flags_s, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv, true); if (rc) return rc; }
    StepConsts c; { int rc = make_consts(g, p, &c); if (rc) return rc; }
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(group_slots && group_base && J && WP && FLG && A_s && Y_s && flags_s, "NULL argument");
    SPHNCA_CHECK_ARG(WB != nullptr || !p->use_alpha, "the blur weights WB are required with an alpha channel");
    const int blocks = (int)((gv.n + 31) / 32);           // one warp per group of 8 particles, 4 groups per CTA
    if (c.use_alpha) SPHNCA_LAUNCH(step_perceive_pairs_kernel<true>, blocks, 128, 0, as_stream(stream_), gv.n, group_slots, group_base, J,
                                   (const float2*)WP, (const float2*)WB, FLG, A_s, c, Y_s, flags_s);
    else SPHNCA_LAUNCH(step_perceive_pairs_kernel<false>, blocks, 128, 0, as_stream(stream_), gv.n, group_slots, group_base, J,
                       (const float2*)WP, (const float2*)WB, FLG, A_s, c, Y_s, flags_s);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_step_mask_pairs(const sphnca_grid* g, const int32_t* group_slots, const int64_t* group_base, const int32_t* J,
                               const float* WB, const uint8_t* alive_in, int32_t mode, uint8_t* flags_s, float* A_new_s,
                               uint8_t* alive_out, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv, true); if (rc) return rc; }
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(group_slots && group_base && J && WB && alive_in && flags_s, "NULL argument");
    SPHNCA_CHECK_ARG(mode == 0 || (mode == 1 && A_new_s != nullptr), "mode 0 (previous mask) or 1 (new mask: needs A_new_s)");
    const float cw = (float)poly6_norm(g->dim, g->h);
    const int blocks = (int)((gv.n + 255) / 256);
    if (mode == 0) SPHNCA_LAUNCH(step_mask_pairs_kernel<0>, blocks, 256, 0, as_stream(stream_), gv.n, group_slots, group_base, J, (const float2*)WB,
                                 alive_in, cw, flags_s, A_new_s, alive_out);
    else SPHNCA_LAUNCH(step_mask_pairs_kernel<1>, blocks, 256, 0, as_stream(stream_), gv.n, group_slots, group_base, J, (const float2*)WB,
                       alive_in, cw, flags_s, A_new_s, alive_out);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_step_perceive_pairs_frame(const sphnca_grid* g, const int32_t* group_slots, const int64_t* group_base, const int32_t* J,
                                         const float* WP, const float* WPZ, const float* WB, const uint8_t* FLG, const float* frame_s,
                                         const float* A_s, const sphnca_step_params* p, float* Y_s, uint8_t* flags_s, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv, true); if (rc) return rc; }
    StepConsts c; { int rc = make_consts(g, p, &c); if (rc) return rc; }
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(group_slots && group_base && J && WP && WPZ && FLG && frame_s && A_s && Y_s && flags_s, "NULL argument");
    SPHNCA_CHECK_ARG(WB != nullptr || !p->use_alpha, "the blur weights WB are required with an alpha channel");
    SPHNCA_CHECK_ARG(((uintptr_t)frame_s & 15) == 0, "frame_s must be 16-byte aligned");
    const int blocks = (int)((gv.n + 31) / 32);
    if (c.use_alpha) SPHNCA_LAUNCH(step_perceive_pairs_frame_kernel<true>, blocks, 128, 0, as_stream(stream_), gv.n, group_slots, group_base, J,
                                   (const float2*)WP, WPZ, (const float2*)WB, FLG, (const float4*)frame_s, A_s, c, Y_s, flags_s);
    else SPHNCA_LAUNCH(step_perceive_pairs_frame_kernel<false>, blocks, 128, 0, as_stream(stream_), gv.n, group_slots, group_base, J,
                       (const float2*)WP, WPZ, (const float2*)WB, FLG, (const float4*)frame_s, A_s, c, Y_s, flags_s);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_step_perceive_pairs_fired(const sphnca_grid* g, const int32_t* rows, const int32_t* count, const int32_t* group_slots,
                                         const int64_t* group_base, const int32_t* J, const float* WP, const float* A_s,
                                         const sphnca_step_params* p, float* Y_c, float* A_c, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv, true); if (rc) return rc; }
    StepConsts c; { int rc = make_consts(g, p, &c); if (rc) return rc; }
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(rows && count && group_slots && group_base && J && WP && A_s && Y_c && A_c, "NULL argument");
    // (with an alpha channel the previous life mask must come from sphnca_step_mask_pairs: this sweep has no mask part)
    const int blocks = (int)((gv.n + 31) / 32);           // sized for all N particles; warps past *count exit at once
    SPHNCA_LAUNCH(step_perceive_pairs_fired_kernel, blocks, 128, 0, as_stream(stream_), rows, count, group_slots, group_base, J,
                  (const float2*)WP, A_s, c, Y_c, A_c);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

static int launch_adjoint_pairs(const sphnca_grid* g, const int32_t* group_slots, const int64_t* group_base, const int32_t* J,
                                const float* WA, const float* SXY, const float* xv4, const float* dY_s, const float* dAdir_s,
                                const sphnca_step_params* p, float* dA_s, const uint8_t* flags_s, const float* dAn_s, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv, true); if (rc) return rc; }
    StepConsts c; { int rc = make_consts(g, p, &c); if (rc) return rc; }
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(group_slots && group_base && J && WA && SXY && xv4 && dY_s && dAdir_s && dA_s, "NULL argument");
    const int blocks = (int)((gv.n + 31) / 32);           // one warp per group of 8 particles, 4 groups per CTA
#define SPHNCA_ADJ_LAUNCH(FIRED_, ILV_)                                                                                                  \
    SPHNCA_LAUNCH((step_adjoint_pairs_kernel<FIRED_, ILV_>), blocks, 128, 0, as_stream(stream_), gv.n, group_slots, group_base, J,     \
                  (const float2*)WA, (const float2*)SXY, (const float4*)xv4, dY_s, dAdir_s, c, dA_s, flags_s, dAn_s)
    if (flags_s != nullptr) { if (c.dy_ilv) SPHNCA_ADJ_LAUNCH(true, true); else SPHNCA_ADJ_LAUNCH(true, false); }
    else { if (c.dy_ilv) SPHNCA_ADJ_LAUNCH(false, true); else SPHNCA_ADJ_LAUNCH(false, false); }
#undef SPHNCA_ADJ_LAUNCH
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_step_adjoint_pairs(const sphnca_grid* g, const int32_t* group_slots, const int64_t* group_base, const int32_t* J,
                                  const float* WA, const float* SXY, const float* xv4, const float* dY_s, const float* dAdir_s,
                                  const sphnca_step_params* p, float* dA_s, void* stream_) {
    return launch_adjoint_pairs(g, group_slots, group_base, J, WA, SXY, xv4, dY_s, dAdir_s, p, dA_s, nullptr, nullptr, stream_);
}

int32_t sphnca_step_adjoint_pairs_fired(const sphnca_grid* g, const int32_t* group_slots, const int64_t* group_base, const int32_t* J,
                                        const float* WA, const float* SXY, const float* xv4, const float* dY_s, const float* dAdir_s,
                                        const uint8_t* flags_s, const float* dA_next_s, const sphnca_step_params* p, float* dA_s,
                                        void* stream_) {
    SPHNCA_CHECK_ARG(flags_s && dA_next_s, "NULL argument");
    return launch_adjoint_pairs(g, group_slots, group_base, J, WA, SXY, xv4, dY_s, dAdir_s, p, dA_s, flags_s, dA_next_s, stream_);
}

}  // extern "C"
