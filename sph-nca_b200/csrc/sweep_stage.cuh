// Block-staged neighbour sweeps: candidates out of shared memory, two passes (in-radius test, then the heavy body).
//
// A CTA owns T consecutive cell-sorted particles.  They lie in a few cell rows (same section, y, z; one row in the
// common case, two when the tile crosses a row end).  For a block row whose particles cover x-cells [xa, xb], the
// candidates of ALL its particles in one of the 3^(D-1) neighbouring rows are ONE contiguous range of the sorted
// arrays, cells [xa-1, xb+1] of that row -- plus, at a domain edge, the wrapped cell (dx-1 for x-cell 0, 0 for x-cell
// dx-1) as a separate segment; axes with dims < 3 fall out of the same rule (the three neighbour cells map onto
// already staged cells, visited as often as the reference visits them).
//
// Prologue: plan the segments (<= MAXROWS x 3^(D-1) x 3), claim shared-memory space, copy every segment with
// coalesced 16-byte loads (all in flight at once: ONE L2 round trip per tile instead of one per first touch inside the
// pair loop -- ncu on the gather kernels: L1 hit rate 75 %, "no eligible warp" 52-60 % of the cycles).  Payload rows
// are stored with a 16-byte pad (row pitch ROWF+4 floats) so that lanes reading DIFFERENT rows hit different banks.
// Pass 1: every thread walks its candidates (positions only) and appends the in-radius ones to its own hit list
// [slot][thread]; pass 2: the heavy body on the hits, in candidate order (sums unchanged -> bit-identical to the
// gather kernels).  A warp runs a loop body as long as ANY lane needs it: testing and accumulating in one loop costs
// ~114 heavy iterations per warp for ~32 neighbours per lane (ncu: 18 of 32 lanes active), the two passes ~114 light
// + ~40 heavy.
// A tile that does not fit (more than MAXROWS rows or more than CAP candidates: tiny grids, very dense cells, 3-D
// clouds) returns false and the caller runs its global-memory loop.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace sphnca {

template <int D, int T, int ROWF, int CAP, int SLOTS, int MAXROWS = 4>
struct StageSmem {                                        // SLOTS == 0: no hit lists (single pass), rows unpadded
    static constexpr int NB = D == 3 ? 9 : 3;             // neighbouring rows per block row
    static constexpr int RP4 = ROWF / 4 + (SLOTS > 0 ? 1 : 0);   // row pitch in float4 (padded for the two-pass variant)
    float4 x[CAP];
    float4 row[CAP * RP4];
    unsigned short list[SLOTS * T + 1];
    // per (block row, neighbouring row): main segment = cells [lo, hi] (global start gs, n candidates, smem offset so),
    // wrapped cells dx-1 (wl) and 0 (wr)
    struct Seg { int gs, so, n, lo, hi, gs_wl, so_wl, n_wl, gs_wr, so_wr, n_wr; };
    Seg seg[MAXROWS * NB];
    int row_id[MAXROWS], xa[MAXROWS], xb[MAXROWS];
    int warp_cnt[T / 32];
    int total;
    uint64_t bar;                                         // completion of the bulk copies (single-pass variant)
};

// Plan the tile: rows, segments, shared-memory offsets.  Returns true when it fits (uniform over the CTA); k = index of
// this thread's block row, nseg = number of plan entries.  Must be called by all T threads of the CTA; `valid` threads
// own the particle at sorted position p in cell `my_cell`.
template <int D, int T, int ROWF, int CAP, int SLOTS, int MAXROWS>
__device__ __forceinline__ bool plan_tile(const GridView& g, const float4* __restrict__ xv4, const float* __restrict__ payload,
                                          StageSmem<D, T, ROWF, CAP, SLOTS, MAXROWS>& sm, int64_t p, bool valid, int my_cell, int& k, int& nseg) {
    using S = StageSmem<D, T, ROWF, CAP, SLOTS, MAXROWS>;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dx = g.dims[0], dy = g.dims[1], dz = D == 3 ? g.dims[2] : 1;
    const int my_row = valid ? my_cell / dx : -1;
    const int cx = valid ? my_cell - my_row * dx : 0;
    // distinct rows of the tile, in order: a particle opens a row when its predecessor in the tile is in another one
    const bool first = valid && (tid == 0 || g.sorted_cell[p - 1] / dx != my_row);
    const bool last = valid && (tid == T - 1 || p + 1 >= g.n || g.sorted_cell[p + 1] / dx != my_row);
    const unsigned ball = __ballot_sync(0xffffffffu, first);
    if (lane == 0) sm.warp_cnt[warp] = __popc(ball);
    if (tid == 0) { sm.total = 0; if (SLOTS == 0) { tc::mbar_init(&sm.bar, 1); tc::mbar_fence_init(); } }
    __syncthreads();
    int before = __popc(ball & ((2u << lane) - 1u));      // rows opened up to and including this lane
    int nrows = 0;
#pragma unroll
    for (int w = 0; w < T / 32; ++w) { const int c = sm.warp_cnt[w]; nrows += c; if (w < warp) before += c; }
    k = before - 1;
    if (nrows > MAXROWS) return false;
    if (first) { sm.row_id[k] = my_row; sm.xa[k] = cx; }
    if (last) sm.xb[k] = cx;
    __syncthreads();
    // plan: one thread per (block row, neighbouring row)
    constexpr int NB = S::NB;
    nseg = nrows * NB;
    if (tid < nseg) {
        typename S::Seg sg{};
        const int kr = tid / NB, o = tid - kr * NB;
        const int oz = D == 3 ? o / 3 : 1, oy = o - (D == 3 ? oz * 3 : 0);
        const int row = sm.row_id[kr];
        const int rows_per_section = dy * dz;
        const int section = row / rows_per_section, rem = row - section * rows_per_section;
        const int cz = D == 3 ? rem / dy : 0, cy = D == 3 ? rem - cz * dy : rem;
        const int z = D == 3 ? (cz + dz + oz - 1) % dz : 0;
        const int y = (cy + dy + oy - 1) % dy;
        const int nrow = (section * rows_per_section + z * dy + y) * dx;          // first global cell of the neighbouring row
        const int xa = sm.xa[kr], xb = sm.xb[kr];
        sg.lo = xa > 0 ? xa - 1 : 0;
        sg.hi = xb < dx - 1 ? xb + 1 : dx - 1;
        sg.gs = g.cell_off[nrow + sg.lo];
        sg.n = g.cell_off[nrow + sg.hi + 1] - sg.gs;
        if (xa == 0 && sg.hi < dx - 1) { sg.gs_wl = g.cell_off[nrow + dx - 1]; sg.n_wl = g.cell_off[nrow + dx] - sg.gs_wl; }
        if (xb == dx - 1 && sg.lo > 0) { sg.gs_wr = g.cell_off[nrow]; sg.n_wr = g.cell_off[nrow + 1] - sg.gs_wr; }
        const int so = atomicAdd(&sm.total, sg.n + sg.n_wl + sg.n_wr);
        sg.so = so; sg.so_wl = so + sg.n; sg.so_wr = so + sg.n + sg.n_wl;
        sm.seg[tid] = sg;
    }
    __syncthreads();
    return sm.total <= CAP;
}

// Plan and fetch.  Returns true when the tile is staged (uniform over the CTA).
template <int D, int T, int ROWF, int CAP, int SLOTS, int MAXROWS>
__device__ __forceinline__ bool stage_tile(const GridView& g, const float4* __restrict__ xv4, const float* __restrict__ payload,
                                           StageSmem<D, T, ROWF, CAP, SLOTS, MAXROWS>& sm, int64_t p, bool valid, int my_cell, int& k) {
    using S = StageSmem<D, T, ROWF, CAP, SLOTS, MAXROWS>;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int nseg = 0;
    if (!plan_tile(g, xv4, payload, sm, p, valid, my_cell, k, nseg)) return false;
    if (SLOTS == 0) {
        // single-pass variant (unpadded rows): every segment is two TMA bulk copies (positions, payload rows) that
        // complete on one mbarrier -- no copy instructions in the threads at all (measured: 0.266 ms vs 0.293 ms with
        // the thread copies below on the padded 2-D configuration)
        const int total = sm.total;
        if (tid == 0) tc::mbar_arrive_expect_tx(&sm.bar, (uint32_t)total * (16u + ROWF * 4u));
        __syncthreads();
        if (tid < nseg) {
            const typename S::Seg sg = sm.seg[tid];
            auto fetch = [&](int gs, int so, int n) {
                if (n > 0) {
                    tc::bulk_g2s(&sm.x[so], xv4 + gs, (uint32_t)n * 16u, &sm.bar);
                    tc::bulk_g2s(&sm.row[(size_t)so * S::RP4], payload + (size_t)gs * ROWF, (uint32_t)n * ROWF * 4u, &sm.bar);
                }
            };
            fetch(sg.gs, sg.so, sg.n);
            fetch(sg.gs_wl, sg.so_wl, sg.n_wl);
            fetch(sg.gs_wr, sg.so_wr, sg.n_wr);
        }
        if (total > 0) tc::mbar_wait(&sm.bar, 0);
        return true;
    }
    // two-pass variant: warps take the plan entries round-robin; lanes run over the 16-byte pieces of a segment
    // (coalesced) and store the rows at the padded pitch
    __syncthreads();
    return true;
}

// fn(j) for every candidate of the particle in cell `my_cell` of block row k, j = index into sm.x / sm.row (row j
// starts at sm.row[j * RP4]); same visiting order as for_each_candidate.
template <int D, int T, int ROWF, int CAP, int SLOTS, int MAXROWS, typename Fn>
__device__ __forceinline__ void for_each_staged(const GridView& g, const StageSmem<D, T, ROWF, CAP, SLOTS, MAXROWS>& sm, int my_cell, int k, Fn&& fn) {
    using S = StageSmem<D, T, ROWF, CAP, SLOTS, MAXROWS>;
    const int dx = g.dims[0], dy = g.dims[1], dz = D == 3 ? g.dims[2] : 1;
    const int row = my_cell / dx, cx = my_cell - row * dx;
    const int rows_per_section = dy * dz;
    const int section = row / rows_per_section, rem = row - section * rows_per_section;
    const int cz = D == 3 ? rem / dy : 0, cy = D == 3 ? rem - cz * dy : rem;
    const int x0 = (cx + dx - 1) % dx, x2 = (cx + 1) % dx;
    const bool contiguous = (x0 + 1 == cx) && (cx + 1 == x2);
    constexpr int NZ = D == 3 ? 3 : 1;
#pragma unroll 1
    for (int oz = 0; oz < NZ; ++oz) {
        const int z = D == 3 ? (cz + dz + oz - 1) % dz : 0;
#pragma unroll 1
        for (int oy = 0; oy < 3; ++oy) {
            const int y = (cy + dy + oy - 1) % dy;
            const int nrow = (section * rows_per_section + z * dy + y) * dx;
            const typename S::Seg& sg = sm.seg[k * S::NB + (D == 3 ? oz * 3 : 0) + oy];
            const int lo = sg.lo, hi = sg.hi, gs = sg.gs, so = sg.so;
            auto run = [&](int xc0, int xc1) {               // cells [xc0, xc1] of the neighbouring row, inside one segment
                const int q0 = g.cell_off[nrow + xc0], q1 = g.cell_off[nrow + xc1 + 1];
                int j = (xc0 >= lo && xc0 <= hi) ? so + (q0 - gs) : (xc0 == dx - 1 ? sg.so_wl : sg.so_wr);
                for (const int je = j + (q1 - q0); j < je; ++j) fn(j);
            };
            if (contiguous) run(x0, x2);
            else { run(x0, x0); run(cx, cx); run(x2, x2); }
        }
    }
}

// The two passes: `near(xj)` = in-radius test on a staged position, `body(j)` = heavy work on a hit.
template <int D, int T, int ROWF, int CAP, int SLOTS, int MAXROWS, typename Near, typename Body>
__device__ __forceinline__ void staged_two_pass(const GridView& g, StageSmem<D, T, ROWF, CAP, SLOTS, MAXROWS>& sm, int my_cell, int k,
                                                Near&& near, Body&& body) {
    unsigned short* mine = sm.list + threadIdx.x;
    int cnt = 0;
    auto flush = [&]() {
        for (int t = 0; t < cnt; ++t) body((int)mine[t * T]);
        cnt = 0;
    };
    for_each_staged(g, sm, my_cell, k, [&](int j) {
        if (near(sm.x[j])) {
            mine[cnt * T] = (unsigned short)j;
            if (++cnt == SLOTS) flush();                   // rare: a full list is worked off in place
        }
    });
    flush();
}

}  // namespace sphnca
