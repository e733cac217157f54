// Neighbour-sweep operators of sphops, general in feature count F and
// in D in {2,3}.  One thread per particle, particles walked in cell-sorted order; all particle
// state is read from the physically cell-sorted arrays (xv4 float4 {x,y,z,v}, row-contiguous
// features), so a warp's candidate reads are a few contiguous, mostly shared (broadcast) runs.
// Results are scattered back to the caller's original order (full rows).
//
// These are the reference-API operators.  The NCA hot loop uses the fused kernels in step.cu.
//
// Arithmetic: r, d^2 in fp32 exactly as the reference (operators_batch.py:72,224; utils dot_nd);
// weights in fp32 with IEEE sqrt/div (the reference mixes fp64 in because h is a Python float);
// `volume` accumulates in fp64 and ends with __frcp_rn like the reference (:77); `count` compares
// in fp64 like the reference (kernels_impl.py:17).  Backward kernels are gather-based adjoints:
// no float atomics anywhere, results are deterministic.
#include "common.cuh"

namespace sphnca {

constexpr int kSweepThreads = 128;

template <int D>
__device__ __forceinline__ float dist2(const float4& a, const float4& b, float& rx, float& ry, float& rz) {
    rx = b.x - a.x; ry = b.y - a.y;
    float d2 = fmaf(ry, ry, rx * rx);
    if (D == 3) { rz = b.z - a.z; d2 = fmaf(rz, rz, d2); } else rz = 0.f;
    return d2;
}

template <int N, bool VEC>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&out)[N]) {
    if (VEC) {
        const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
        for (int k = 0; k < N / 4; ++k) {
            float4 t = __ldg(p4 + k);
            out[4 * k] = t.x; out[4 * k + 1] = t.y; out[4 * k + 2] = t.z; out[4 * k + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) out[k] = __ldg(p + k);
    }
}

// ---- volume --------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kSweepThreads) volume_fwd_kernel(GridView g, const float4* __restrict__ xv4,
                                                                   double h2, double norm, float* __restrict__ v,
                                                                   float4* __restrict__ xv4_out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.n) return;
    const float4 xi = xv4[p];
    double sum = 0.0;
    for_each_candidate<D>(g, g.sorted_cell[p], [&](int q) {
        float rx, ry, rz;
        float4 xj = __ldg(&xv4[q]);
        if (g.periodic) xj = nearest_image(g, xi, xj);
        double t = h2 - (double)dist2<D>(xi, xj, rx, ry, rz);
        sum += t > 0.0 ? t * t * t : 0.0;
    });
    const float vi = __frcp_rn((float)(norm * sum));
    v[g.sorted_id[p]] = vi;
    xv4_out[p] = make_float4(xi.x, xi.y, xi.z, vi);
}

// Volumes of the ghost copies of a periodic plane (sphnca_grid.image_volume).  The copy of particle i shifted by s = (sx, sy)
// periods sits at x_i + s*period in the ghost-padded cloud; its neighbours there are the own particles and the EXISTING copies
// within h: for a periodic neighbour k of i, seen from i at shift t, that is the copy s' = s + t of k (the particle itself when
// s' = 0), if it exists.  A particle lies in at most one band per axis, so it has at most 3 copies: (bx,0), (0,by), (bx,by).
template <int D>
__global__ void __launch_bounds__(kSweepThreads) volume_images_kernel(GridView g, const float4* __restrict__ xv4, double h2, double norm,
                                                                      float* __restrict__ vimg) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.n) return;
    const float4 xi = xv4[p];
    const int bx = xi.x <= g.band_lo[0] ? 1 : (xi.x >= g.band_hi[0] ? -1 : 0);
    const int by = xi.y <= g.band_lo[1] ? 1 : (xi.y >= g.band_hi[1] ? -1 : 0);
    const int isx[3] = {bx, 0, bx}, isy[3] = {0, by, by};
    const bool have[3] = {bx != 0, by != 0, bx != 0 && by != 0};
    double sum[3] = {0.0, 0.0, 0.0};
    if (have[0] || have[1]) {
        for_each_candidate<D>(g, g.sorted_cell[p], [&](int q) {
            const float4 xk = __ldg(&xv4[q]);
            const float4 xk_img = nearest_image(g, xi, xk);
            int tx, ty;
            image_shift(xk, xk_img, tx, ty);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (!have[c]) continue;
                const int sx = isx[c] + tx, sy = isy[c] + ty;          // the copy of k next to copy c of i
                if (sx < -1 || sx > 1 || sy < -1 || sy > 1 || !image_exists(g, xk, sx, sy)) continue;
                // positions as the ghost-padded cloud holds them: one fp32 add of the shift each (nca.py:214-217)
                float4 pi = xi, pk = xk;
                if (isx[c]) pi.x += (float)isx[c] * g.period[0];
                if (isy[c]) pi.y += (float)isy[c] * g.period[1];
                if (sx) pk.x += (float)sx * g.period[0];
                if (sy) pk.y += (float)sy * g.period[1];
                float rx, ry, rz;
                const double t = h2 - (double)dist2<D>(pi, pk, rx, ry, rz);
                sum[c] += t > 0.0 ? t * t * t : 0.0;
            }
        });
    }
    float* o = vimg + (size_t)p * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) o[k] = 0.f;
    o[4] = xi.w;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        if (have[c]) o[(isx[c] + 1) * 3 + (isy[c] + 1)] = __frcp_rn((float)(norm * sum[c]));
}

// dL/dx_i = Cw sum_j (v_i^2 g_i + v_j^2 g_j) * (-6 (h^2-d^2)^2) * r      (operators_batch.py:137-150)
template <int D>
__global__ void __launch_bounds__(kSweepThreads) volume_bwd_kernel(GridView g, const float4* __restrict__ xv4,
                                                                   const float* __restrict__ dLdv_s, float h2,
                                                                   float norm, float* __restrict__ dLdx) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.n) return;
    const float4 xi = xv4[p];
    const float ci = xi.w * xi.w * dLdv_s[p];
    float ax = 0.f, ay = 0.f, az = 0.f;
    for_each_candidate<D>(g, g.sorted_cell[p], [&](int q) {
        float rx, ry, rz;
        const float4 xj = __ldg(&xv4[q]);
        const float d2 = dist2<D>(xi, xj, rx, ry, rz);
        if (d2 < h2) {
            const float t = h2 - d2;
            const float mag = (ci + xj.w * xj.w * __ldg(&dLdv_s[q])) * -6.f * t * t;
            ax = fmaf(mag, rx, ax); ay = fmaf(mag, ry, ay); az = fmaf(mag, rz, az);
        }
    });
    float* o = dLdx + (size_t)g.sorted_id[p] * D;
    o[0] = norm * ax; o[1] = norm * ay;
    if (D == 3) o[2] = norm * az;
}

// ---- count ---------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kSweepThreads) count_fwd_kernel(GridView g, const float4* __restrict__ xv4,
                                                                  double h2, float* __restrict__ c) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.n) return;
    const float4 xi = xv4[p];
    int cnt = 0;
    for_each_candidate<D>(g, g.sorted_cell[p], [&](int q) {
        float rx, ry, rz;
        cnt += (double)dist2<D>(xi, __ldg(&xv4[q]), rx, ry, rz) < h2;
    });
    c[g.sorted_id[p]] = (float)cnt;
}

// ---- gradient ------------------------------------------------------------------------------
// GA_i[f,d] = Cg sum_j (A_j[f]-A_i[f]) v_j m_ij r[d]       (operators_batch.py:221-234)
template <int D, int FC, bool VEC>
__global__ void __launch_bounds__(kSweepThreads) gradient_fwd_kernel(GridView g, const float4* __restrict__ xv4,
                                                                     const float* __restrict__ A_s, int F, int f0,
                                                                     float h, float h2, float norm,
                                                                     float* __restrict__ GA) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.n) return;
    const float4 xi = xv4[p];
    float Ai[FC], acc[FC][D];
    load_row<FC, VEC>(A_s + (size_t)p * F + f0, Ai);
#pragma unroll
    for (int f = 0; f < FC; ++f)
#pragma unroll
        for (int d = 0; d < D; ++d) acc[f][d] = 0.f;
    for_each_candidate<D>(g, g.sorted_cell[p], [&](int q) {
        float r[3];
        const float4 xj = __ldg(&xv4[q]);
        const float d2 = dist2<D>(xi, xj, r[0], r[1], r[2]);
        const float m = spiky_mag(d2, h, h2) * xj.w;
        if (m != 0.f) {
            float Aj[FC];
            load_row<FC, VEC>(A_s + (size_t)q * F + f0, Aj);
            float w[D];
#pragma unroll
            for (int d = 0; d < D; ++d) w[d] = m * r[d];
#pragma unroll
            for (int f = 0; f < FC; ++f) {
                const float dA = Aj[f] - Ai[f];
#pragma unroll
                for (int d = 0; d < D; ++d) acc[f][d] = fmaf(dA, w[d], acc[f][d]);
            }
        }
    });
    float* o = GA + ((size_t)g.sorted_id[p] * F + f0) * D;
#pragma unroll
    for (int f = 0; f < FC; ++f)
#pragma unroll
        for (int d = 0; d < D; ++d) o[f * D + d] = norm * acc[f][d];
}

// dL/dA_i[f] = -Cg sum_j m_ij r . (v_i G_j[f,:] + v_j G_i[f,:])          (operators_batch.py:319-342)
//            = -Cg ( v_i sum_j m_ij r.G_j[f,:]  +  G_i[f,:] . sum_j m_ij v_j r )
template <int D, int FC, bool VEC>
__global__ void __launch_bounds__(kSweepThreads) gradient_bwd_kernel(GridView g, const float4* __restrict__ xv4,
                                                                     const float* __restrict__ G_s, int F, int f0,
                                                                     float h, float h2, float norm,
                                                                     float* __restrict__ dLdA) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.n) return;
    const float4 xi = xv4[p];
    float acc[FC], s[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int f = 0; f < FC; ++f) acc[f] = 0.f;
    for_each_candidate<D>(g, g.sorted_cell[p], [&](int q) {
        float r[3];
        const float4 xj = __ldg(&xv4[q]);
        const float d2 = dist2<D>(xi, xj, r[0], r[1], r[2]);
        const float m = spiky_mag(d2, h, h2);
        if (m != 0.f) {
            float Gj[FC * D];
            load_row<FC * D, VEC>(G_s + ((size_t)q * F + f0) * D, Gj);
            float w[D];
#pragma unroll
            for (int d = 0; d < D; ++d) { w[d] = m * r[d]; s[d] = fmaf(w[d], xj.w, s[d]); }
#pragma unroll
            for (int f = 0; f < FC; ++f)
#pragma unroll
                for (int d = 0; d < D; ++d) acc[f] = fmaf(w[d], Gj[f * D + d], acc[f]);
        }
    });
    float Gi[FC * D];
    load_row<FC * D, VEC>(G_s + ((size_t)p * F + f0) * D, Gi);
    float* o = dLdA + (size_t)g.sorted_id[p] * F + f0;
#pragma unroll
    for (int f = 0; f < FC; ++f) {
        float t = xi.w * acc[f];
#pragma unroll
        for (int d = 0; d < D; ++d) t = fmaf(Gi[f * D + d], s[d], t);
        o[f] = -norm * t;
    }
}

// ---- blur ----------------------------------------------------------------------------------
// SA_i[f] = Cw sum_j A_j[f] W_ij v_j   (operators_batch.py:661-671).  ADJ: the adjoint w.r.t. A,
// dL/dA_i[f] = Cw v_i sum_j W_ij G_j[f]  (W and the neighbour relation are symmetric).
template <int D, int FC, bool VEC, bool ADJ>
__global__ void __launch_bounds__(kSweepThreads) blur_kernel(GridView g, const float4* __restrict__ xv4,
                                                             const float* __restrict__ A_s, int F, int f0, float h2,
                                                             float norm, float* __restrict__ SA) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.n) return;
    const float4 xi = xv4[p];
    float acc[FC];
#pragma unroll
    for (int f = 0; f < FC; ++f) acc[f] = 0.f;
    for_each_candidate<D>(g, g.sorted_cell[p], [&](int q) {
        float rx, ry, rz;
        const float4 xj = __ldg(&xv4[q]);
        const float d2 = dist2<D>(xi, xj, rx, ry, rz);
        if (d2 < h2) {
            const float t = h2 - d2;
            const float w = ADJ ? t * t * t : t * t * t * xj.w;
            float Aj[FC];
            load_row<FC, VEC>(A_s + (size_t)q * F + f0, Aj);
#pragma unroll
            for (int f = 0; f < FC; ++f) acc[f] = fmaf(Aj[f], w, acc[f]);
        }
    });
    const float scale = ADJ ? norm * xi.w : norm;
    float* o = SA + (size_t)g.sorted_id[p] * F + f0;
#pragma unroll
    for (int f = 0; f < FC; ++f) o[f] = acc[f] * scale;
}

// ---- divergence ----------------------------------------------------------------------------
// DA_i[f] = Cg sum_j v_j m_ij (A_j[f,:]-A_i[f,:]) . r                     (operators_batch.py:411-423)
template <int D, int FC>
__global__ void __launch_bounds__(kSweepThreads) divergence_fwd_kernel(GridView g, const float4* __restrict__ xv4,
                                                                       const float* __restrict__ A_s, int F, int f0,
                                                                       float h, float h2, float norm,
                                                                       float* __restrict__ DA) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.n) return;
    const float4 xi = xv4[p];
    float Ai[FC * D], acc[FC];
    load_row<FC * D, false>(A_s + ((size_t)p * F + f0) * D, Ai);
#pragma unroll
    for (int f = 0; f < FC; ++f) acc[f] = 0.f;
    for_each_candidate<D>(g, g.sorted_cell[p], [&](int q) {
        float r[3];
        const float4 xj = __ldg(&xv4[q]);
        const float d2 = dist2<D>(xi, xj, r[0], r[1], r[2]);
        const float m = spiky_mag(d2, h, h2) * xj.w;
        if (m != 0.f) {
            float Aj[FC * D];
            load_row<FC * D, false>(A_s + ((size_t)q * F + f0) * D, Aj);
#pragma unroll
            for (int f = 0; f < FC; ++f) {
                float dot = 0.f;
#pragma unroll
                for (int d = 0; d < D; ++d) dot = fmaf(Aj[f * D + d] - Ai[f * D + d], r[d], dot);
                acc[f] = fmaf(m, dot, acc[f]);
            }
        }
    });
    float* o = DA + (size_t)g.sorted_id[p] * F + f0;
#pragma unroll
    for (int f = 0; f < FC; ++f) o[f] = norm * acc[f];
}

// true adjoint w.r.t. A:  dL/dA_k[f,d] = -Cg ( v_k sum_j m_kj r[d] G_j[f]  +  G_k[f] sum_j m_kj v_j r[d] )
template <int D, int FC>
__global__ void __launch_bounds__(kSweepThreads) divergence_bwd_kernel(GridView g, const float4* __restrict__ xv4,
                                                                       const float* __restrict__ G_s, int F, int f0,
                                                                       float h, float h2, float norm,
                                                                       float* __restrict__ dLdA) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.n) return;
    const float4 xi = xv4[p];
    float T[FC][D], s[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int f = 0; f < FC; ++f)
#pragma unroll
        for (int d = 0; d < D; ++d) T[f][d] = 0.f;
    for_each_candidate<D>(g, g.sorted_cell[p], [&](int q) {
        float r[3];
        const float4 xj = __ldg(&xv4[q]);
        const float d2 = dist2<D>(xi, xj, r[0], r[1], r[2]);
        const float m = spiky_mag(d2, h, h2);
        if (m != 0.f) {
            float Gj[FC];
            load_row<FC, false>(G_s + (size_t)q * F + f0, Gj);
            float w[D];
#pragma unroll
            for (int d = 0; d < D; ++d) { w[d] = m * r[d]; s[d] = fmaf(w[d], xj.w, s[d]); }
#pragma unroll
            for (int f = 0; f < FC; ++f)
#pragma unroll
                for (int d = 0; d < D; ++d) T[f][d] = fmaf(w[d], Gj[f], T[f][d]);
        }
    });
    float Gi[FC];
    load_row<FC, false>(G_s + (size_t)p * F + f0, Gi);
    float* o = dLdA + ((size_t)g.sorted_id[p] * F + f0) * D;
#pragma unroll
    for (int f = 0; f < FC; ++f)
#pragma unroll
        for (int d = 0; d < D; ++d) o[f * D + d] = -norm * fmaf(xi.w, T[f][d], Gi[f] * s[d]);
}

// ---- dispatch ------------------------------------------------------------------------------
static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }
static inline int sweep_blocks(const GridView& g) { return (int)((g.n + kSweepThreads - 1) / kSweepThreads); }

// F is covered greedily by compile-time chunks (the reference chunks by MAX_FEATURES=64, func.py:120-139)
template <typename Launch>
static void for_each_chunk(int F, Launch&& launch) {
    int f0 = 0;
    while (f0 < F) {
        int rem = F - f0;
        int fc = rem >= 16 ? 16 : rem >= 8 ? 8 : rem >= 4 ? 4 : rem >= 2 ? 2 : 1;
        launch(f0, fc);
        f0 += fc;
    }
}

#define SPHNCA_DISPATCH_FC(fc, ...)                    \
    switch (fc) {                                      \
        case 16: { constexpr int FC = 16; __VA_ARGS__; } break; \
        case 8:  { constexpr int FC = 8;  __VA_ARGS__; } break; \
        case 4:  { constexpr int FC = 4;  __VA_ARGS__; } break; \
        case 2:  { constexpr int FC = 2;  __VA_ARGS__; } break; \
        default: { constexpr int FC = 1;  __VA_ARGS__; } break; \
    }

}  // namespace sphnca

using namespace sphnca;

#define SPHNCA_PROLOGUE() SPHNCA_PROLOGUE_P(false)
#define SPHNCA_PROLOGUE_P(allow_periodic)                  \
    GridView gv;                                           \
    { int rc_ = make_view(g, &gv, allow_periodic); if (rc_) return rc_; }  \
    cudaStream_t stream = as_stream(stream_);              \
    if (gv.n == 0) return SPHNCA_OK;                       \
    const double hd = g->h;                                \
    const float h = (float)hd, h2 = (float)(hd * hd);      \
    (void)h; (void)h2;                                     \
    const int blocks = sweep_blocks(gv);

extern "C" {

int32_t sphnca_volume_fwd(const sphnca_grid* g, const float* xv4, float* v, float* xv4_out, void* stream_) {
    SPHNCA_PROLOGUE_P(true);
    SPHNCA_CHECK_ARG(xv4 && v && xv4_out, "NULL argument");
    const double norm = poly6_norm(gv.dim, hd);
    if (gv.dim == 3) SPHNCA_LAUNCH(volume_fwd_kernel<3>, blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, hd * hd, norm, v, (float4*)xv4_out);
    else SPHNCA_LAUNCH(volume_fwd_kernel<2>, blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, hd * hd, norm, v, (float4*)xv4_out);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_volume_images(const sphnca_grid* g, const float* xv4, float* image_volume, void* stream_) {
    SPHNCA_PROLOGUE_P(true);
    SPHNCA_CHECK_ARG(xv4 && image_volume, "NULL argument");
    SPHNCA_CHECK_ARG(gv.period[0] > 0.f && gv.period[1] > 0.f && gv.period[2] == 0.f, "ghost volumes are defined for planes periodic in x and y");
    const double norm = poly6_norm(gv.dim, hd);
    if (gv.dim == 3) SPHNCA_LAUNCH(volume_images_kernel<3>, blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, hd * hd, norm, image_volume);
    else SPHNCA_LAUNCH(volume_images_kernel<2>, blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, hd * hd, norm, image_volume);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_volume_bwd(const sphnca_grid* g, const float* xv4, const float* dLdv_s, float* dLdx, void* stream_) {
    SPHNCA_PROLOGUE();
    SPHNCA_CHECK_ARG(xv4 && dLdv_s && dLdx, "NULL argument");
    const float norm = (float)poly6_norm(gv.dim, hd);
    if (gv.dim == 3) SPHNCA_LAUNCH(volume_bwd_kernel<3>, blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, dLdv_s, h2, norm, dLdx);
    else SPHNCA_LAUNCH(volume_bwd_kernel<2>, blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, dLdv_s, h2, norm, dLdx);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_count_fwd(const sphnca_grid* g, const float* xv4, float* c, void* stream_) {
    SPHNCA_PROLOGUE();
    SPHNCA_CHECK_ARG(xv4 && c, "NULL argument");
    if (gv.dim == 3) SPHNCA_LAUNCH(count_fwd_kernel<3>, blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, hd * hd, c);
    else SPHNCA_LAUNCH(count_fwd_kernel<2>, blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, hd * hd, c);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_gradient_fwd(const sphnca_grid* g, const float* xv4, const float* A_s, int32_t F, float* GA, void* stream_) {
    SPHNCA_PROLOGUE();
    SPHNCA_CHECK_ARG(xv4 && A_s && GA && F > 0, "bad argument");
    const float norm = (float)spiky_norm(gv.dim, hd);
    const bool vec_ok = (F % 4 == 0) && aligned16(A_s);
    for_each_chunk(F, [&](int f0, int fc) {
        const bool vec = vec_ok && fc >= 4;
        SPHNCA_DISPATCH_FC(fc,
            if (gv.dim == 3) {
                if (vec) SPHNCA_LAUNCH((gradient_fwd_kernel<3, FC, (FC >= 4)>), blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, A_s, F, f0, h, h2, norm, GA);
                else SPHNCA_LAUNCH((gradient_fwd_kernel<3, FC, false>), blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, A_s, F, f0, h, h2, norm, GA);
            } else {
                if (vec) SPHNCA_LAUNCH((gradient_fwd_kernel<2, FC, (FC >= 4)>), blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, A_s, F, f0, h, h2, norm, GA);
                else SPHNCA_LAUNCH((gradient_fwd_kernel<2, FC, false>), blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, A_s, F, f0, h, h2, norm, GA);
            })
    });
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_gradient_bwd(const sphnca_grid* g, const float* xv4, const float* G_s, int32_t F, float* dLdA, void* stream_) {
    SPHNCA_PROLOGUE();
    SPHNCA_CHECK_ARG(xv4 && G_s && dLdA && F > 0, "bad argument");
    const float norm = (float)spiky_norm(gv.dim, hd);
    const bool vec_ok = ((F * gv.dim) % 4 == 0) && aligned16(G_s);
    for_each_chunk(F, [&](int f0, int fc) {
        const bool vec = vec_ok && ((f0 * gv.dim) % 4 == 0) && ((fc * gv.dim) % 4 == 0);
        SPHNCA_DISPATCH_FC(fc,
            if (gv.dim == 3) {
                if (vec) SPHNCA_LAUNCH((gradient_bwd_kernel<3, FC, (FC * 3 % 4 == 0)>), blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, G_s, F, f0, h, h2, norm, dLdA);
                else SPHNCA_LAUNCH((gradient_bwd_kernel<3, FC, false>), blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, G_s, F, f0, h, h2, norm, dLdA);
            } else {
                if (vec) SPHNCA_LAUNCH((gradient_bwd_kernel<2, FC, (FC * 2 % 4 == 0)>), blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, G_s, F, f0, h, h2, norm, dLdA);
                else SPHNCA_LAUNCH((gradient_bwd_kernel<2, FC, false>), blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, G_s, F, f0, h, h2, norm, dLdA);
            })
    });
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

}  // extern "C"

template <bool ADJ>
static int32_t blur_impl(const sphnca_grid* g, const float* xv4, const float* A_s, int32_t F, float* SA, void* stream_) {
    SPHNCA_PROLOGUE();
    SPHNCA_CHECK_ARG(xv4 && A_s && SA && F > 0, "bad argument");
    const float norm = (float)poly6_norm(gv.dim, hd);
    const bool vec_ok = (F % 4 == 0) && aligned16(A_s);
    for_each_chunk(F, [&](int f0, int fc) {
        const bool vec = vec_ok && fc >= 4;
        SPHNCA_DISPATCH_FC(fc,
            if (gv.dim == 3) {
                if (vec) SPHNCA_LAUNCH((blur_kernel<3, FC, (FC >= 4), ADJ>), blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, A_s, F, f0, h2, norm, SA);
                else SPHNCA_LAUNCH((blur_kernel<3, FC, false, ADJ>), blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, A_s, F, f0, h2, norm, SA);
            } else {
                if (vec) SPHNCA_LAUNCH((blur_kernel<2, FC, (FC >= 4), ADJ>), blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, A_s, F, f0, h2, norm, SA);
                else SPHNCA_LAUNCH((blur_kernel<2, FC, false, ADJ>), blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, A_s, F, f0, h2, norm, SA);
            })
    });
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}
extern "C" {

int32_t sphnca_blur_fwd(const sphnca_grid* g, const float* xv4, const float* A_s, int32_t F, float* SA, void* s) {
    return blur_impl<false>(g, xv4, A_s, F, SA, s);
}
int32_t sphnca_blur_bwd(const sphnca_grid* g, const float* xv4, const float* G_s, int32_t F, float* dLdA, void* s) {
    return blur_impl<true>(g, xv4, G_s, F, dLdA, s);
}

int32_t sphnca_divergence_fwd(const sphnca_grid* g, const float* xv4, const float* A_s, int32_t F, float* DA, void* stream_) {
    SPHNCA_PROLOGUE();
    SPHNCA_CHECK_ARG(xv4 && A_s && DA && F > 0, "bad argument");
    const float norm = (float)spiky_norm(gv.dim, hd);
    // rows are F*D wide: chunks of at most 8 features keep the register tile small
    int f0 = 0;
    while (f0 < F) {
        int rem = F - f0;
        int fc = rem >= 8 ? 8 : rem >= 4 ? 4 : rem >= 2 ? 2 : 1;
        SPHNCA_DISPATCH_FC(fc,
            if (FC <= 8) {
                if (gv.dim == 3) SPHNCA_LAUNCH((divergence_fwd_kernel<3, (FC <= 8 ? FC : 8)>), blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, A_s, F, f0, h, h2, norm, DA);
                else SPHNCA_LAUNCH((divergence_fwd_kernel<2, (FC <= 8 ? FC : 8)>), blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, A_s, F, f0, h, h2, norm, DA);
            })
        f0 += fc;
    }
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_divergence_bwd(const sphnca_grid* g, const float* xv4, const float* G_s, int32_t F, float* dLdA, void* stream_) {
    SPHNCA_PROLOGUE();
    SPHNCA_CHECK_ARG(xv4 && G_s && dLdA && F > 0, "bad argument");
    const float norm = (float)spiky_norm(gv.dim, hd);
    int f0 = 0;
    while (f0 < F) {
        int rem = F - f0;
        int fc = rem >= 8 ? 8 : rem >= 4 ? 4 : rem >= 2 ? 2 : 1;
        SPHNCA_DISPATCH_FC(fc,
            if (FC <= 8) {
                if (gv.dim == 3) SPHNCA_LAUNCH((divergence_bwd_kernel<3, (FC <= 8 ? FC : 8)>), blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, G_s, F, f0, h, h2, norm, dLdA);
                else SPHNCA_LAUNCH((divergence_bwd_kernel<2, (FC <= 8 ? FC : 8)>), blocks, kSweepThreads, 0, stream, gv, (const float4*)xv4, G_s, F, f0, h, h2, norm, dLdA);
            })
        f0 += fc;
    }
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

}  // extern "C"
