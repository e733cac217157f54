// Shared device/host helpers for libsphnca (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sphnca.h"

namespace sphnca {

void set_error(const char* fmt, ...);
void count_launches(int n);   // bookkeeping for sphnca_launch_count()

#define SPHNCA_CHECK_ARG(cond, ...)        \
    do {                                   \
        if (!(cond)) {                     \
            sphnca::set_error(__VA_ARGS__);\
            return SPHNCA_ERR_INVALID;     \
        }                                  \
    } while (0)

#define SPHNCA_CUDA(call)                                                                   \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            sphnca::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return SPHNCA_ERR_CUDA;                                                         \
        }                                                                                   \
    } while (0)

#define SPHNCA_LAUNCH_CHECK() SPHNCA_CUDA(cudaGetLastError())
// every kernel launch of the library goes through this macro so that launches can be counted
#define SPHNCA_LAUNCH(kernel, grid, block, smem, stream, ...) \
    do { kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); sphnca::count_launches(1); } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr float kPi = 3.14159265358979323846f;

// Device-side copy of the cell tables (passed by value to kernels).
struct GridView {
    int dim;
    int num_sections;
    int64_t n;
    int cells;        // cells per section
    int dims[3];
    const int32_t* __restrict__ cell_off;
    const int32_t* __restrict__ sorted_id;
    const int32_t* __restrict__ sorted_cell;
    float period[3];  // > 0: periodic along that axis (sphnca_grid.period)
    int periodic;
    const float* __restrict__ image_volume;   // ghost-volume emulation (sphnca_grid.image_volume), may be NULL
    float band_lo[2], band_hi[2];
};

// Does the ghost copy of the particle at b, shifted by (sx, sy) periods, exist in the reference's ghost-padded cloud?
__device__ __forceinline__ bool image_exists(const GridView& g, const float4& b, int sx, int sy) {
    const bool okx = sx == 0 || (sx == 1 ? b.x <= g.band_lo[0] : b.x >= g.band_hi[0]);
    const bool oky = sy == 0 || (sy == 1 ? b.y <= g.band_lo[1] : b.y >= g.band_hi[1]);
    return okx && oky;
}
// shift (in periods, -1 / 0 / +1 per axis) that nearest_image applied to b
__device__ __forceinline__ void image_shift(const float4& b, const float4& b_img, int& sx, int& sy) {
    sx = b_img.x > b.x ? 1 : (b_img.x < b.x ? -1 : 0);
    sy = b_img.y > b.y ? 1 : (b_img.y < b.y ? -1 : 0);
}

// The neighbour's position as the reference's ghost copy would hold it (nca.py:214-217: x_w = x + (+-gsize), one fp32 add): the image
// of b closest to a along every periodic axis.  Open axes (period 0) are left alone.
__device__ __forceinline__ float4 nearest_image(const GridView& g, const float4& a, float4 b) {
    if (g.period[0] > 0.f) { const float d = b.x - a.x; if (d > 0.5f * g.period[0]) b.x -= g.period[0]; else if (d < -0.5f * g.period[0]) b.x += g.period[0]; }
    if (g.period[1] > 0.f) { const float d = b.y - a.y; if (d > 0.5f * g.period[1]) b.y -= g.period[1]; else if (d < -0.5f * g.period[1]) b.y += g.period[1]; }
    if (g.period[2] > 0.f) { const float d = b.z - a.z; if (d > 0.5f * g.period[2]) b.z -= g.period[2]; else if (d < -0.5f * g.period[2]) b.z += g.period[2]; }
    return b;
}

// allow_periodic: only the entries that implement the periodic neighbour search (or never look at positions) pass true
inline int make_view(const sphnca_grid* g, GridView* v, bool allow_periodic = false) {
    if (!g) { set_error("grid is NULL"); return SPHNCA_ERR_INVALID; }
    if (g->dim != 2 && g->dim != 3) { set_error("grid.dim must be 2 or 3, got %d", g->dim); return SPHNCA_ERR_INVALID; }
    if (g->num_particles < 0 || g->num_particles >= (int64_t(1) << 31)) { set_error("num_particles out of range"); return SPHNCA_ERR_INVALID; }
    if (g->cells_per_section <= 0 || g->cells_per_section * g->num_sections >= (int64_t(1) << 31)) {
        set_error("cell table too large for int32 indexing"); return SPHNCA_ERR_INVALID;
    }
    v->dim = g->dim; v->num_sections = g->num_sections; v->n = g->num_particles;
    v->cells = (int)g->cells_per_section;
    for (int d = 0; d < 3; ++d) v->dims[d] = d < g->dim ? (int)g->dims[d] : 1;
    v->cell_off = g->cell_off; v->sorted_id = g->sorted_id; v->sorted_cell = g->sorted_cell;
    v->periodic = 0;
    for (int d = 0; d < 3; ++d) {
        v->period[d] = (d < g->dim && g->period[d] > 0.0) ? (float)g->period[d] : 0.f;
        if (v->period[d] > 0.f) {
            v->periodic = 1;
            // the modulo wrap of the cell hash must BE the period (and a stencil must not meet itself)
            const double cells = g->period[d] / g->h;
            if (g->dims[d] < 3 || cells < (double)g->dims[d] - 1e-6 * g->dims[d] || cells > (double)g->dims[d] + 1e-6 * g->dims[d]) {
                set_error("periodic axis %d: period %g must equal dims (%lld) x h (%g), dims >= 3", d, g->period[d], (long long)g->dims[d], g->h);
                return SPHNCA_ERR_INVALID;
            }
        }
    }
    v->image_volume = v->periodic ? g->image_volume : nullptr;
    for (int d = 0; d < 2; ++d) { v->band_lo[d] = (float)g->band_lo[d]; v->band_hi[d] = (float)g->band_hi[d]; }
    if (v->image_volume != nullptr && v->period[2] > 0.f) { set_error("ghost-volume emulation is for periodic PLANES (x, y)"); return SPHNCA_ERR_INVALID; }
    if (v->periodic && !allow_periodic) { set_error("this entry does not implement periodic grids (sphnca_grid.period)"); return SPHNCA_ERR_INVALID; }
    return SPHNCA_OK;
}

// Normalisation constants, reference kernels_impl.py:20-30 (poly6) and :131-141 (spiky).
// Evaluated in double on the host like the reference's Python floats, then rounded once.
inline double poly6_norm(int dim, double h) {
    const double pi = 3.14159265358979323846;
    double h2 = h * h, h4 = h2 * h2, h8 = h4 * h4;
    return dim == 2 ? 4.0 / (pi * h8) : 315.0 / (64.0 * pi * h8 * h);
}
inline double spiky_norm(int dim, double h) {
    const double pi = 3.14159265358979323846;
    double h2 = h * h, h4 = h2 * h2;
    return dim == 2 ? 10.0 / (pi * h4 * h) : 15.0 / (pi * h4 * h2);
}

// ------------------------------------------------------------------------------------------
// Neighbour enumeration.  For the particle at sorted position p the reference visits the 3^D
// cells around its own cell, offsets (-1,0,1) per axis with x fastest, each wrapped modulo dims
// (single_impl.py:46-60, operators_batch.py:57-76).  An axis with dims<3 therefore visits the
// same cell more than once -- reproduced here because the loop is the same.
// `fn(q)` is called for every candidate sorted position q, in the reference's order.
// ------------------------------------------------------------------------------------------
// `fn(q0, q1)` is called for every contiguous run [q0,q1) of candidate sorted positions, in the
// reference's visiting order.  The three x-cells of a stencil row are one run when they do not wrap.
template <int D, typename Fn>
__device__ __forceinline__ void for_each_run(const GridView& g, int gcell, Fn&& fn) {
    const int section = gcell / g.cells;
    const int cell = gcell - section * g.cells;
    const int base = section * g.cells;
    const int dx = g.dims[0], dy = g.dims[1];
    int cx, cy, cz = 0;
    if (D == 3) {
        cz = cell / (dx * dy);
        int rem = cell - cz * dx * dy;
        cy = rem / dx; cx = rem - cy * dx;
    } else {
        cy = cell / dx; cx = cell - cy * dx;
    }
    // wrapped neighbour coordinates per axis: (c + dims + d) % dims for d = -1, 0, 1
    const int x0 = (cx + dx - 1) % dx, x2 = (cx + 1) % dx;
    const bool contiguous = (x0 + 1 == cx) && (cx + 1 == x2);
    constexpr int NZ = D == 3 ? 3 : 1;
#pragma unroll 1
    for (int oz = 0; oz < NZ; ++oz) {
        const int z = D == 3 ? (cz + g.dims[2] + oz - 1) % g.dims[2] : 0;
#pragma unroll 1
        for (int oy = 0; oy < 3; ++oy) {
            const int y = (cy + dy + oy - 1) % dy;
            const int row = base + (z * dy + y) * dx;
            if (contiguous) {
                fn(g.cell_off[row + x0], g.cell_off[row + x2 + 1]);
            } else {
                fn(g.cell_off[row + x0], g.cell_off[row + x0 + 1]);
                fn(g.cell_off[row + cx], g.cell_off[row + cx + 1]);
                fn(g.cell_off[row + x2], g.cell_off[row + x2 + 1]);
            }
        }
    }
}

// `fn(q)` for every candidate sorted position q, in the reference's order.
template <int D, typename Fn>
__device__ __forceinline__ void for_each_candidate(const GridView& g, int gcell, Fn&& fn) {
    for_each_run<D>(g, gcell, [&](int q0, int q1) {
        for (int q = q0; q < q1; ++q) fn(q);
    });
}

// In-radius compaction.  Only ~35% (padded 2-D) / ~15% (3-D) of the candidates of a 3^D-cell sweep lie
// inside the support radius, but a warp executes a branch body whenever ANY lane takes it, so a heavy body
// (16-channel FMAs) would run for every candidate.  Two passes instead: a light pass tests every candidate
// and appends the in-radius sorted positions to a per-thread list in shared memory (layout [slot][thread]:
// conflict-free), a heavy pass `fn(q, xj)` walks only the list (trip count = the lane's neighbour count).
// A full list is flushed in place, so any neighbour count works.  Visiting order is unchanged.
constexpr int kNbrSlots = 64;
template <int D, int THREADS, typename Fn>
__device__ __forceinline__ void for_each_neighbor(const GridView& g, int gcell, const float4* __restrict__ xv4,
                                                  const float4& xi, float h2, int* __restrict__ list, Fn&& fn) {
    int* mine = list + threadIdx.x;
    int cnt = 0;
    auto flush = [&]() {
        for (int k = 0; k < cnt; ++k) {
            const int q = mine[k * THREADS];
            fn(q, __ldg(&xv4[q]));
        }
        cnt = 0;
    };
    for_each_run<D>(g, gcell, [&](int q0, int q1) {
        for (int q = q0; q < q1; ++q) {
            const float4 xj = __ldg(&xv4[q]);
            const float rx = xj.x - xi.x, ry = xj.y - xi.y;
            float d2 = fmaf(ry, ry, rx * rx);
            if (D == 3) { const float rz = xj.z - xi.z; d2 = fmaf(rz, rz, d2); }
            if (d2 < h2) {
                mine[cnt * THREADS] = q;
                if (++cnt == kNbrSlots) flush();
            }
        }
    });
    flush();
}

// Packed fp32 pairs (Blackwell FFMA2 / FADD2): one instruction performs two IEEE fp32 operations, so the
// per-neighbour channel loops (16 channels) issue half the instructions.  Results are bit-identical to the
// scalar fmaf / subtraction on each half.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long&>(d))
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
          "l"(reinterpret_cast<unsigned long long&>(c)));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(d))
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return d;
}
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) {
    float2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(d))
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return d;
}

// spiky gradient weight 3 (h-d)^2 / d for 0 < d < h, else 0 (kernels_impl.py:143-156).
// The reference returns exact 0 for r == 0 (`r[i] and mag*r[i]`), also for coincident particles.
__device__ __forceinline__ float spiky_mag(float d2, float h, float h2) {
    if (!(d2 < h2) || d2 <= 0.f) return 0.f;
    const float d = sqrtf(d2);
    const float hd = h - d;
    return 3.f * hd * hd / d;
}

// poly6 smoothing (h^2 - d^2)^3 clamped at 0 (kernels_impl.py:32-39)
__device__ __forceinline__ float poly6_w(float d2, float h2) {
    const float t = h2 - d2;
    return t > 0.f ? t * t * t : 0.f;
}

}  // namespace sphnca
