// Hash-grid neighbour build for all sections of a packed batch in one pass.
//
// Replaces, per batch element, the reference's grid_hash_kernel + torch.sort + cell_index_init_kernel
// (+2 device syncs each; code/sphops/preprocess.py:33-41, hashgrid/single_impl.py:62-124).
//
// B200 design: the key range is known (B*cells), so the "radix sort by cell" is a counting sort:
//   K1 hash_count   1 thread/particle: cell hash (exact f64 floor-divide, Python modulo) + int histogram
//   K2 scan         3-phase exclusive scan of the histogram -> cell_off (+ reference cell_start/cell_end)
//   K3 scatter      claim a slot inside the cell with an int atomic (order inside the cell arbitrary)
//   K4 rank         order each cell's members by original index -> deterministic, identical to a
//                   STABLE sort by cell hash (what torch.sort on CUDA gives the reference)
// Only integer atomics are used; the result is bit-reproducible.  Everything is HBM-bound integer
// work: ~12 B read + 3x4 B written per particle in K1/K3/K4, 3 passes over the (B*cells) table in K2.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace sphnca {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

static unsigned long long g_launches = 0;
void count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
long long launches() { return (long long)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;                       // per thread
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int find_section(const int32_t* __restrict__ sec_off, int B, int n) {
    int lo = 0, hi = B;                              // largest b with sec_off[b] <= n
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (sec_off[mid] <= n) lo = mid; else hi = mid;
    }
    return lo;
}

// K1.  hi[n] = section*cells + sum_d c_d * (floor(x_d/h) mod dims_d)   (single_impl.py:34-44)
template <int D>
__global__ void __launch_bounds__(256) hash_count_kernel(const float* __restrict__ x, int n_total,
                                                          const int32_t* __restrict__ sec_off, int B, double h,
                                                          int d0, int d1, int d2, int cells,
                                                          int32_t* __restrict__ hi, int32_t* __restrict__ cnt) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_total) return;
    const int dims[3] = {d0, d1, d2};
    long long out = 0, c = 1;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        // int(math.floor(x / h)) with x promoted to double: IEEE division, then floor
        long long gi = (long long)floor((double)x[(size_t)n * D + d] / h);
        long long m = gi % dims[d];
        if (m < 0) m += dims[d];                     // Python modulo: sign of the divisor
        out += c * m;
        c *= dims[d];
    }
    int b = find_section(sec_off, B, n);
    int gcell = b * cells + (int)out;
    hi[n] = gcell;
    atomicAdd(&cnt[gcell], 1);
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const int32_t* __restrict__ cnt, int m,
                                                                   int32_t* __restrict__ block_sums) {
    __shared__ int warp_sums[kScanThreads / 32];
    int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) if (base + i < m) s += cnt[base + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = threadIdx.x < kScanThreads / 32 ? warp_sums[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = v;
    }
}

// single block: exclusive scan of block_sums in place
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(int32_t* __restrict__ block_sums, int nb) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int start = 0; start < nb; start += 1024) {
        int i = start + threadIdx.x;
        int v = i < nb ? block_sums[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o) inc += t;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warp_tot[threadIdx.x];
            int winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, winc, o);
                if (threadIdx.x >= o) winc += t;
            }
            warp_tot[threadIdx.x] = winc - w;        // exclusive over warps
        }
        __syncthreads();
        int carry = carry_s;
        int excl = carry + warp_tot[threadIdx.x >> 5] + inc - v;
        if (i < nb) block_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
}

// K2c.  cell_off (global, exclusive) + the reference's section-local cell_start / cell_end (-1 = empty,
// single_impl.py:93-124).
__global__ void __launch_bounds__(kScanThreads) scan_write_kernel(const int32_t* __restrict__ cnt, int m,
                                                                  const int32_t* __restrict__ block_sums,
                                                                  const int32_t* __restrict__ sec_off, int cells,
                                                                  int n_total, int32_t* __restrict__ cell_off,
                                                                  int32_t* __restrict__ cell_start,
                                                                  int32_t* __restrict__ cell_end) {
    __shared__ int warp_tot[kScanThreads / 32];
    int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int vals[kScanItems];
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) { vals[i] = base + i < m ? cnt[base + i] : 0; s += vals[i]; }
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
        int w = threadIdx.x < kScanThreads / 32 ? warp_tot[threadIdx.x] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (threadIdx.x >= o) winc += t;
        }
        if (threadIdx.x < kScanThreads / 32) warp_tot[threadIdx.x] = winc - w;
    }
    __syncthreads();
    int run = block_sums[blockIdx.x] + warp_tot[threadIdx.x >> 5] + inc - s;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        int c = base + i;
        if (c < m) {
            cell_off[c] = run;
            int local = run - sec_off[c / cells];
            cell_start[c] = vals[i] > 0 ? local : -1;
            cell_end[c] = vals[i] > 0 ? local + vals[i] : -1;
            run += vals[i];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cell_off[m] = n_total;
}

// K3.  cnt counts down to zero: slot inside the cell is arbitrary, fixed up by K4.
__global__ void __launch_bounds__(256) scatter_kernel(const int32_t* __restrict__ hi, int n_total,
                                                      const int32_t* __restrict__ cell_off,
                                                      int32_t* __restrict__ cnt, int32_t* __restrict__ tmp) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_total) return;
    int c = hi[n];
    int slot = atomicSub(&cnt[c], 1) - 1;
    tmp[cell_off[c] + slot] = n;
}

// K4.  rank of n among its cell's members = number of members with a smaller original index.
__global__ void __launch_bounds__(256) rank_kernel(const int32_t* __restrict__ hi, const int32_t* __restrict__ tmp,
                                                   int n_total, const int32_t* __restrict__ cell_off,
                                                   const int32_t* __restrict__ sec_off, int cells,
                                                   int32_t* __restrict__ sorted_id, int32_t* __restrict__ sorted_cell,
                                                   int64_t* __restrict__ idx) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_total) return;
    int n = tmp[q];
    int c = hi[n];
    int lo = cell_off[c], end = cell_off[c + 1];
    int rank = 0;
    for (int k = lo; k < end; ++k) rank += tmp[k] < n;
    int pos = lo + rank;
    sorted_id[pos] = n;
    sorted_cell[pos] = c;
    idx[pos] = (int64_t)(n - sec_off[c / cells]);
}

__global__ void num_points_kernel(const int32_t* __restrict__ sec_off, int B, int32_t* __restrict__ num_points) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) num_points[b] = sec_off[b + 1] - sec_off[b];
}

// ---- layout helpers ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_xv_kernel(const float* __restrict__ x, const float* __restrict__ v,
                                                      const int32_t* __restrict__ sorted_id, int n_total, int D,
                                                      float4* __restrict__ xv4) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_total) return;
    int i = sorted_id[p];
    float4 o;
    o.x = x[(size_t)i * D]; o.y = x[(size_t)i * D + 1]; o.z = D == 3 ? x[(size_t)i * D + 2] : 0.f;
    o.w = v ? v[i] : 0.f;
    xv4[p] = o;
}

// one thread per float (row-contiguous): coalesced on the sorted side, row-granular on the other
template <bool PACK>
__global__ void __launch_bounds__(256) permute_rows_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                           const int32_t* __restrict__ sorted_id, int64_t total,
                                                           int row) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int64_t p = t / row;
    int f = (int)(t - p * row);
    int64_t i = sorted_id[p];
    if (PACK) dst[t] = src[i * row + f];
    else dst[i * row + f] = src[t];
}
template <bool PACK>
__global__ void __launch_bounds__(256) permute_rows4_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                                            const int32_t* __restrict__ sorted_id, int64_t total,
                                                            int row4) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int64_t p = t / row4;
    int f = (int)(t - p * row4);
    int64_t i = sorted_id[p];
    if (PACK) dst[t] = src[i * row4 + f];
    else dst[i * row4 + f] = src[t];
}

__global__ void __launch_bounds__(256) pack_bytes_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                         const int32_t* __restrict__ sorted_id, int n_total) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_total) dst[p] = src[sorted_id[p]];
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct BuildWorkspace {
    int32_t *hi, *tmp, *cnt, *block_sums, *sec_off;
    size_t bytes;
};
static BuildWorkspace carve(void* base, int64_t N, int B, int64_t M) {
    BuildWorkspace w;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off = align_up(off + n, 256); return o; };
    int64_t nb = (M + kScanTile - 1) / kScanTile;
    size_t o_hi = take(sizeof(int32_t) * (size_t)(N > 0 ? N : 1));
    size_t o_tmp = take(sizeof(int32_t) * (size_t)(N > 0 ? N : 1));
    size_t o_cnt = take(sizeof(int32_t) * (size_t)M);
    size_t o_bs = take(sizeof(int32_t) * (size_t)(nb > 0 ? nb : 1));
    size_t o_so = take(sizeof(int32_t) * (size_t)(B + 1));
    char* b = reinterpret_cast<char*>(base);
    w.hi = (int32_t*)(b + o_hi); w.tmp = (int32_t*)(b + o_tmp); w.cnt = (int32_t*)(b + o_cnt);
    w.block_sums = (int32_t*)(b + o_bs); w.sec_off = (int32_t*)(b + o_so);
    w.bytes = off;
    return w;
}

}  // namespace sphnca

using namespace sphnca;

extern "C" {

int32_t sphnca_abi_version(void) { return 3; }
const char* sphnca_last_error(void) { return sphnca::last_error(); }
int64_t sphnca_launch_count(void) { return sphnca::launches(); }

int32_t sphnca_hashgrid_workspace_bytes(int64_t N, int32_t B, int64_t cells, size_t* bytes) {
    SPHNCA_CHECK_ARG(bytes != nullptr, "bytes is NULL");
    SPHNCA_CHECK_ARG(N >= 0 && B >= 0 && cells > 0, "bad sizes N=%lld B=%d cells=%lld", (long long)N, B, (long long)cells);
    SPHNCA_CHECK_ARG((int64_t)B * cells < (int64_t(1) << 31) - kScanTile, "B*cells does not fit int32");
    *bytes = carve(nullptr, N, B, (int64_t)B * cells).bytes;
    return SPHNCA_OK;
}

int32_t sphnca_hashgrid_build(const float* x, int64_t N, int32_t D, const int64_t* sections, int32_t B, double h,
                              const int64_t* dims, int64_t* idx, int32_t* cell_start, int32_t* cell_end,
                              int32_t* num_points, int32_t* cell_off, int32_t* sorted_id, int32_t* sorted_cell,
                              void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = as_stream(stream_);
    SPHNCA_CHECK_ARG(D == 2 || D == 3, "D must be 2 or 3, got %d", D);
    SPHNCA_CHECK_ARG(B >= 0 && N >= 0 && N < (int64_t(1) << 31), "bad N/B");
    SPHNCA_CHECK_ARG(sections && dims, "sections/dims is NULL");
    SPHNCA_CHECK_ARG(h > 0.0, "h must be positive");
    int64_t cells = 1, total = 0;
    for (int d = 0; d < D; ++d) { SPHNCA_CHECK_ARG(dims[d] > 0, "dims[%d] must be positive", d); cells *= dims[d]; }
    for (int b = 0; b < B; ++b) { SPHNCA_CHECK_ARG(sections[b] >= 0, "negative section size"); total += sections[b]; }
    SPHNCA_CHECK_ARG(total == N, "sum(sections)=%lld != N=%lld", (long long)total, (long long)N);
    const int64_t M = (int64_t)B * cells;
    SPHNCA_CHECK_ARG(M < (int64_t(1) << 31) - kScanTile, "B*cells does not fit int32");
    if (B == 0) return SPHNCA_OK;
    BuildWorkspace w = carve(workspace, N, B, M);
    if (workspace_bytes < w.bytes || !workspace) {
        set_error("hashgrid workspace too small: %zu < %zu", workspace_bytes, w.bytes);
        return SPHNCA_ERR_WORKSPACE;
    }
    // section offsets -> device (tiny; pageable-source async copy is staged before returning)
    int32_t sec_host[1025];
    int32_t* sec = B + 1 <= 1025 ? sec_host : (int32_t*)malloc(sizeof(int32_t) * (B + 1));
    sec[0] = 0;
    for (int b = 0; b < B; ++b) sec[b + 1] = sec[b] + (int32_t)sections[b];
    cudaError_t ce = cudaMemcpyAsync(w.sec_off, sec, sizeof(int32_t) * (B + 1), cudaMemcpyHostToDevice, stream);
    if (sec != sec_host) { cudaStreamSynchronize(stream); free(sec); }
    SPHNCA_CUDA(ce);
    SPHNCA_CUDA(cudaMemsetAsync(w.cnt, 0, sizeof(int32_t) * (size_t)M, stream));

    const int nN = (int)((N + 255) / 256);
    const int nb = (int)((M + kScanTile - 1) / kScanTile);
    SPHNCA_LAUNCH(num_points_kernel, (B + 127) / 128, 128, 0, stream, w.sec_off, B, num_points);
    if (N > 0) {
        if (D == 3)
            SPHNCA_LAUNCH(hash_count_kernel<3>, nN, 256, 0, stream, x, (int)N, w.sec_off, B, h, (int)dims[0], (int)dims[1], (int)dims[2], (int)cells, w.hi, w.cnt);
        else
            SPHNCA_LAUNCH(hash_count_kernel<2>, nN, 256, 0, stream, x, (int)N, w.sec_off, B, h, (int)dims[0], (int)dims[1], 1, (int)cells, w.hi, w.cnt);
    }
    SPHNCA_LAUNCH(scan_reduce_kernel, nb, kScanThreads, 0, stream, w.cnt, (int)M, w.block_sums);
    SPHNCA_LAUNCH(scan_block_sums_kernel, 1, 1024, 0, stream, w.block_sums, nb);
    SPHNCA_LAUNCH(scan_write_kernel, nb, kScanThreads, 0, stream, w.cnt, (int)M, w.block_sums, w.sec_off, (int)cells, (int)N,
                                                       cell_off, cell_start, cell_end);
    if (N > 0) {
        SPHNCA_LAUNCH(scatter_kernel, nN, 256, 0, stream, w.hi, (int)N, cell_off, w.cnt, w.tmp);
        SPHNCA_LAUNCH(rank_kernel, nN, 256, 0, stream, w.hi, w.tmp, (int)N, cell_off, w.sec_off, (int)cells, sorted_id, sorted_cell, idx);
    }
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_pack_xv(const sphnca_grid* g, const float* x, const float* v, float* xv4, void* stream) {
    GridView gv; int rc = make_view(g, &gv, true); if (rc) return rc;
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(x && xv4, "x/xv4 is NULL");
    SPHNCA_LAUNCH(pack_xv_kernel, (int)((gv.n + 255) / 256), 256, 0, as_stream(stream), x, v, gv.sorted_id, (int)gv.n, gv.dim, (float4*)xv4);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

static int32_t permute_rows(const sphnca_grid* g, const float* src, int32_t row, float* dst, void* stream, bool pack) {
    GridView gv; int rc = make_view(g, &gv, true); if (rc) return rc;
    SPHNCA_CHECK_ARG(row > 0, "row must be positive");
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(src && dst, "src/dst is NULL");
    const bool vec = (row % 4 == 0) && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
    if (vec) {
        int64_t total = gv.n * (row / 4);
        int blocks = (int)((total + 255) / 256);
        if (pack) SPHNCA_LAUNCH(permute_rows4_kernel<true>, blocks, 256, 0, as_stream(stream), (const float4*)src, (float4*)dst, gv.sorted_id, total, row / 4);
        else SPHNCA_LAUNCH(permute_rows4_kernel<false>, blocks, 256, 0, as_stream(stream), (const float4*)src, (float4*)dst, gv.sorted_id, total, row / 4);
    } else {
        int64_t total = gv.n * row;
        int64_t blocks = (total + 255) / 256;
        SPHNCA_CHECK_ARG(blocks < (int64_t(1) << 31), "too many elements");
        if (pack) SPHNCA_LAUNCH(permute_rows_kernel<true>, (unsigned)blocks, 256, 0, as_stream(stream), src, dst, gv.sorted_id, total, row);
        else SPHNCA_LAUNCH(permute_rows_kernel<false>, (unsigned)blocks, 256, 0, as_stream(stream), src, dst, gv.sorted_id, total, row);
    }
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_pack_rows(const sphnca_grid* g, const float* src, int32_t row, float* dst, void* stream) {
    return permute_rows(g, src, row, dst, stream, true);
}
int32_t sphnca_unpack_rows(const sphnca_grid* g, const float* src, int32_t row, float* dst, void* stream) {
    return permute_rows(g, src, row, dst, stream, false);
}
int32_t sphnca_pack_bytes(const sphnca_grid* g, const uint8_t* src, uint8_t* dst, void* stream) {
    GridView gv; int rc = make_view(g, &gv, true); if (rc) return rc;
    if (gv.n == 0) return SPHNCA_OK;
    SPHNCA_LAUNCH(pack_bytes_kernel, (int)((gv.n + 255) / 256), 256, 0, as_stream(stream), src, dst, gv.sorted_id, (int)gv.n);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

}  // extern "C"
