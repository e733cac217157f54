// Tile sweep engine: neighbour sweeps out of shared memory (used by the hot step kernels).
//
// A CTA owns T consecutive cell-sorted particles.  They are processed per SEGMENT = the particles of the
// tile that share one cell row (same section, y, z; usually one segment, two when the tile crosses a row end).
// For a segment covering x-cells [xs, xe] and each of the 3^(D-1) neighbouring rows, the candidates of all its
// particles are ONE contiguous run of the sorted arrays, cells [xs-1, xe+1] of that row -- plus, at a domain
// edge, the wrapped cell as a separate run (cell dx-1 for particles in x-cell 0, cell 0 for x-cell dx-1;
// this also reproduces the double/triple counting of the reference on axes with dims < 3).
// Each run is staged in chunks of CAP candidates with cp.async (positions float4 + ROW payload floats at a
// padded stride), then every thread
//   pass 1  walks the part of the chunk that belongs to ITS three x-cells, tests d^2 < h^2 and appends the
//           hits to a per-thread list in shared memory ([slot][thread], 2-byte indices),
//   pass 2  runs the heavy `body(xj, row)` only on the hits (trip count = that lane's neighbours in the chunk).
// Only ~35% (padded 2-D) / ~15% (3-D) of the candidates are hits, and a warp executes a branch body whenever
// ANY lane takes it -- without the two passes the heavy body would run for every candidate.
// All neighbour data is read from shared memory (no L1/L2 round trips in the loops); global traffic is
// contiguous and asynchronous.  Visiting order differs from the reference only in fp32 summation order.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace sphnca {

template <int T, int ROW, int CAP = 128, int SLOTS = 48>
struct TileSweepSmem {
    static constexpr int ROWP = ROW + 4;                 // padded row stride (floats): conflict-free LDS.128
    float4 x[CAP];
    float row[CAP * ROWP];
    unsigned short list[SLOTS * T];
    int seg_key;                                         // broadcast slots
    int seg_xe;
};

template <int D, int T, int ROW, int CAP, int SLOTS, typename Body>
__device__ __forceinline__ void tile_sweep(const GridView& g, const float4* __restrict__ xv4,
                                           const float* __restrict__ payload, TileSweepSmem<T, ROW, CAP, SLOTS>& sm,
                                           int p0, float h2, const float4& xi, bool valid, int my_cell, Body&& body) {
    using S = TileSweepSmem<T, ROW, CAP, SLOTS>;
    const int tid = threadIdx.x;
    const int dx = g.dims[0], dy = g.dims[1], dz = D == 3 ? g.dims[2] : 1;
    const int n_tile = (int)((g.n - p0) < T ? (g.n - p0) : T);
    // my cell coordinates
    const int section = valid ? my_cell / g.cells : 0;
    const int cell = valid ? my_cell - section * g.cells : 0;
    const int row_key = valid ? my_cell / dx : -1;       // section, z, y
    const int cx = cell % dx;
    int seg_begin = 0;
    while (seg_begin < n_tile) {
        if (tid == seg_begin) sm.seg_key = row_key;
        __syncthreads();
        const int key = sm.seg_key;
        const bool in_seg = valid && row_key == key;
        const int seg_cnt = __syncthreads_count(in_seg);
        if (tid == seg_begin + seg_cnt - 1) sm.seg_xe = cx;
        // segment geometry (uniform): section / row coordinates from the key, x range from first / last particle
        const int s_sec = key / (dy * dz);
        const int rem = key - s_sec * (dy * dz);
        const int cz = D == 3 ? rem / dy : 0;
        const int cy = D == 3 ? rem - cz * dy : rem;
        const int base = s_sec * g.cells;
        __syncthreads();
        const int xe = sm.seg_xe;
        int xs;
        {   // x-cell of the first particle of the segment (uniform value, read by everyone from global: L1 hit)
            const int c0 = g.sorted_cell[p0 + seg_begin];
            xs = (c0 - (c0 / g.cells) * g.cells) % dx;
        }
        constexpr int NZ = D == 3 ? 3 : 1;
#pragma unroll 1
        for (int oz = 0; oz < NZ; ++oz) {
            const int z = D == 3 ? (cz + dz + oz - 1) % dz : 0;
#pragma unroll 1
            for (int oy = 0; oy < 3; ++oy) {
                const int y = (cy + dy + oy - 1) % dy;
                const int nrow = base + (z * dy + y) * dx;
                // run 0: in-range cells [max(xs-1,0), min(xe+1,dx-1)]; run 1: wrapped cell dx-1 (for cx == 0);
                // run 2: wrapped cell 0 (for cx == dx-1)
#pragma unroll 1
                for (int run = 0; run < 3; ++run) {
                    int ca, cb;                          // cell range [ca, cb] of this run
                    if (run == 0) { ca = xs > 0 ? xs - 1 : 0; cb = xe < dx - 1 ? xe + 1 : dx - 1; }
                    else if (run == 1) { if (xs != 0) continue; ca = cb = dx - 1; }
                    else { if (xe != dx - 1) continue; ca = cb = 0; }
                    const int q0 = g.cell_off[nrow + ca], q1 = g.cell_off[nrow + cb + 1];
                    if (q0 >= q1) continue;              // empty run (uniform)
                    // my sub-range of this run
                    int lo = 0, hi = 0;
                    if (in_seg) {
                        if (run == 0) {
                            const int a = cx > 0 ? cx - 1 : 0, b = cx < dx - 1 ? cx + 1 : dx - 1;
                            lo = g.cell_off[nrow + a]; hi = g.cell_off[nrow + b + 1];
                        } else if ((run == 1 && cx == 0) || (run == 2 && cx == dx - 1)) { lo = q0; hi = q1; }
                    }
#pragma unroll 1
                    for (int c0 = q0; c0 < q1; c0 += CAP) {
                        const int cn = (q1 - c0) < CAP ? (q1 - c0) : CAP;
                        // ---- stage chunk [c0, c0+cn): positions + payload rows, 16-byte cp.async pieces ----
                        constexpr int PIECES = 1 + ROW / 4;
                        for (int i = tid; i < cn * PIECES; i += T) {
                            const int j = i / PIECES, k = i - j * PIECES;
                            const void* src; void* dst;
                            if (k == 0) { src = xv4 + (c0 + j); dst = &sm.x[j]; }
                            else { src = payload + (size_t)(c0 + j) * ROW + 4 * (k - 1); dst = &sm.row[j * S::ROWP + 4 * (k - 1)]; }
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(tc::smem_u32(dst)), "l"(src) : "memory");
                        }
                        asm volatile("cp.async.commit_group;" ::: "memory");
                        asm volatile("cp.async.wait_group 0;" ::: "memory");
                        __syncthreads();
                        // ---- pass 1: my candidates inside the chunk -> hit list ----
                        const int ja = (lo > c0 ? lo : c0) - c0;
                        const int jb = (hi < c0 + cn ? hi : c0 + cn) - c0;
                        int cnt = 0;
                        unsigned short* mine = sm.list + tid;
                        for (int j = ja; j < jb; ++j) {
                            const float4 xj = sm.x[j];
                            const float rx = xj.x - xi.x, ry = xj.y - xi.y;
                            float d2 = fmaf(ry, ry, rx * rx);
                            if (D == 3) { const float rz = xj.z - xi.z; d2 = fmaf(rz, rz, d2); }
                            if (d2 < h2) {
                                mine[cnt * T] = (unsigned short)j;
                                if (++cnt == SLOTS) {    // rare: flush a full list in place
                                    for (int k = 0; k < cnt; ++k) { const int jj = mine[k * T]; body(sm.x[jj], &sm.row[jj * S::ROWP]); }
                                    cnt = 0;
                                }
                            }
                        }
                        // ---- pass 2: heavy body on the hits ----
                        for (int k = 0; k < cnt; ++k) { const int jj = mine[k * T]; body(sm.x[jj], &sm.row[jj * S::ROWP]); }
                        __syncthreads();                 // chunk buffers are reused
                    }
                }
            }
        }
        seg_begin += seg_cnt;
    }
}

}  // namespace sphnca
