// Fire lists: the MLP half of an NCA step only matters for the particles whose fire-mask draw came up.
//
// Reference code/nca.py:100-107 evaluates the MLP for EVERY particle and then keeps the old state where the
// stochastic update mask is off (`torch.where(update_mask, nA, A)`): for those particles the MLP output is
// discarded in the forward, and in the backward no gradient reaches the MLP (dL/dO == 0, so they add exactly 0
// to dW1, db1, dW2, db2 and to dL/dY).  With the default fire rate 0.5 that is half of all tensor-core work.
//
//   fire_list        rows[0..count) = sorted positions of the fired particles, in increasing order (ordered
//                    compaction: block counts -> single-block scan -> in-block ranks; no atomics, deterministic);
//                    the not-fired rows get their new state right here: A_new = A, alive byte from A; optionally the
//                    static life-mask bit of every row (no alpha channel: pairs.cu FLG) is copied into the flag bytes.
//   bwd_unfired      the backward of those rows: dL/dY = 0, direct dL/dA = (living ? dL/dA_new : 0).
//
// The tensor-core kernels (step_mlp_tc.cu, step_bwd_tc.cu, wgrad_tc.cu) then run over the `count` listed rows only;
// `count` stays on the device (they read it from memory), so there is no host synchronisation.
#include "step_common.cuh"

namespace sphnca {

namespace fl {
constexpr int THREADS = 256, PER_THREAD = 4, BLOCK_ROWS = THREADS * PER_THREAD;   // 1024 rows per block
}

__global__ void __launch_bounds__(fl::THREADS)
fire_count_kernel(int64_t n, const uint8_t* __restrict__ fire_s, int32_t* __restrict__ block_count) {
    using namespace fl;
    const int64_t base = (int64_t)blockIdx.x * BLOCK_ROWS + (int64_t)threadIdx.x * PER_THREAD;
    int c = 0;
    if (base + PER_THREAD <= n) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(fire_s + base);      // 4 bytes, 4-aligned (base % 4 == 0)
        c = ((w & 0xffu) != 0) + ((w & 0xff00u) != 0) + ((w & 0xff0000u) != 0) + ((w & 0xff000000u) != 0);
    } else {
        for (int k = 0; k < PER_THREAD; ++k) c += (base + k < n && fire_s[base + k] != 0) ? 1 : 0;
    }
    __shared__ int wsum[THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) t += wsum[w];
        block_count[blockIdx.x] = t;
    }
}

// The draw itself, fused with the count: fire byte of sorted row p = (u[sorted_id[p]] <= rate), u = the reference's
// `torch.rand(N)` in particle order (nca.py:100) -- replaces torch's compare, the byte permutation and fire_count_kernel.
__global__ void __launch_bounds__(fl::THREADS)
fire_draw_count_kernel(int64_t n, const float* __restrict__ u, float rate, const int32_t* __restrict__ sorted_id,
                       uint8_t* __restrict__ fire_s, int32_t* __restrict__ block_count) {
    using namespace fl;
    const int64_t base = (int64_t)blockIdx.x * BLOCK_ROWS + (int64_t)threadIdx.x * PER_THREAD;
    int c = 0;
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) {
        const int64_t p = base + k;
        const uint32_t f = (p < n && __ldg(u + __ldg(sorted_id + p)) <= rate) ? 1u : 0u;
        packed |= f << (8 * k); c += (int)f;
    }
    if (base + PER_THREAD <= n) *reinterpret_cast<uint32_t*>(fire_s + base) = packed;
    else for (int k = 0; k < PER_THREAD; ++k) if (base + k < n) fire_s[base + k] = (uint8_t)((packed >> (8 * k)) & 1u);
    __shared__ int wsum[THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) t += wsum[w];
        block_count[blockIdx.x] = t;
    }
}

// exclusive scan of the block counts in place (one block; fixed order), total -> count[0]
__global__ void __launch_bounds__(1024)
fire_scan_kernel(int nb, int32_t* __restrict__ block_count, int32_t* __restrict__ count) {
    __shared__ int wsum[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const int v = i < nb ? block_count[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = wsum[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
            wsum[threadIdx.x] = w;
        }
        __syncthreads();
        const int carry = carry_s;
        const int incl = x + (threadIdx.x >= 32 ? wsum[(threadIdx.x >> 5) - 1] : 0);
        if (i < nb) block_count[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) count[0] = carry_s;
}

__global__ void __launch_bounds__(fl::THREADS)
fire_fill_kernel(int64_t n, const uint8_t* __restrict__ fire_s, const int32_t* __restrict__ block_off,
                 const float* __restrict__ A_s, int use_alpha, int32_t* __restrict__ rows,
                 float* __restrict__ A_new_s, uint8_t* __restrict__ alive_s, const uint8_t* __restrict__ FLG,
                 uint8_t* __restrict__ flags_s, int fold_mask, const uint8_t* __restrict__ carry_sel) {
    using namespace fl;
    const int64_t base = (int64_t)blockIdx.x * BLOCK_ROWS + (int64_t)threadIdx.x * PER_THREAD;
    bool f[PER_THREAD];
    int c = 0;
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) { f[k] = base + k < n && fire_s[base + k] != 0; c += f[k] ? 1 : 0; }
    // exclusive rank of this thread's first row among the block's fired rows
    __shared__ int wsum[THREADS / 32];
    int x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    int before = x - c;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += wsum[w];
    int slot = block_off[blockIdx.x] + before;
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) {
        const int64_t p = base + k;
        if (p >= n) break;
        // static life mask (no alpha channel): the previous-mask bit of every row; fold_mask: the new mask is the same bit
        // (step_mask_static_kernel), so bit 1 is set here too and dead rows are zeroed below / by the MLP kernel
        const bool live = FLG == nullptr || FLG[p] != 0;
        if (FLG != nullptr && flags_s != nullptr) flags_s[p] = live ? (fold_mask ? 3 : 1) : 0;
        if (f[k]) { rows[slot++] = (int32_t)p; continue; }
        if (A_new_s == nullptr) continue;                   // list only
        if (carry_sel != nullptr && carry_sel[p] == 0) continue;   // this row is carried over by the other list of the step (multigpu.py)
        // not fired: the state is carried over unchanged (reference nca.py:107); fire bit of the flags stays 0
        const float4* a4 = reinterpret_cast<const float4*>(A_s + p * NF);
        float4* o4 = reinterpret_cast<float4*>(A_new_s + p * NF);
        const float4 a0 = __ldg(a4), a1 = __ldg(a4 + 1), a2 = __ldg(a4 + 2), a3 = __ldg(a4 + 3);
        if (fold_mask && !live) {                           // the product with 0 of step_mask_static_kernel, bit for bit
            o4[0] = make_float4(a0.x * 0.f, a0.y * 0.f, a0.z * 0.f, a0.w * 0.f); o4[1] = make_float4(a1.x * 0.f, a1.y * 0.f, a1.z * 0.f, a1.w * 0.f);
            o4[2] = make_float4(a2.x * 0.f, a2.y * 0.f, a2.z * 0.f, a2.w * 0.f); o4[3] = make_float4(a3.x * 0.f, a3.y * 0.f, a3.z * 0.f, a3.w * 0.f);
        } else { o4[0] = a0; o4[1] = a1; o4[2] = a2; o4[3] = a3; }
        alive_s[p] = use_alpha ? (a0.w > 0.1f ? 1 : 0) : 1;
    }
}

// Counter-based fire bytes for slab rollouts (multigpu.py): fire byte of sorted row p = gid[p] >= 0 && U(gid[p], seed) <= rate with
// U = the top 24 bits of splitmix64(gid + golden * (seed + 1)) -- multigpu.counter_uniform(ids, 1, seed), bit for bit -- so
// the draw of a particle depends on its GLOBAL id only (any rank count).  Also emits the two masked copies the boundary /
// interior passes need.  One thread = 4 rows.
__global__ void __launch_bounds__(256)
fire_bytes_counter_kernel(int64_t n, const int64_t* __restrict__ gid_s, uint64_t seed_term, float rate,
                          const uint8_t* __restrict__ bnd_s, uint8_t* __restrict__ fire_s, uint8_t* __restrict__ fire_b,
                          uint8_t* __restrict__ fire_i) {
    const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (base >= n) return;
    uint32_t f = 0, fb = 0, fi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t p = base + k;
        if (p >= n) break;
        const int64_t g = gid_s[p];
        uint64_t z = (uint64_t)g + seed_term;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z = z ^ (z >> 31);
        const float u = (float)(uint32_t)((z >> 40) & 0xFFFFFFu) * (1.0f / 16777216.0f);
        const uint32_t fire = (g >= 0 && u <= rate) ? 1u : 0u;
        const uint32_t b = bnd_s != nullptr ? (uint32_t)bnd_s[p] : 0u;
        f |= fire << (8 * k); fb |= (fire & b) << (8 * k); fi |= (fire & (b ^ 1u)) << (8 * k);
    }
    if (base + 4 <= n) {
        *reinterpret_cast<uint32_t*>(fire_s + base) = f;
        if (fire_b != nullptr) { *reinterpret_cast<uint32_t*>(fire_b + base) = fb; *reinterpret_cast<uint32_t*>(fire_i + base) = fi; }
    } else {
        for (int k = 0; k < 4 && base + k < n; ++k) {
            fire_s[base + k] = (uint8_t)((f >> (8 * k)) & 1u);
            if (fire_b != nullptr) { fire_b[base + k] = (uint8_t)((fb >> (8 * k)) & 1u); fire_i[base + k] = (uint8_t)((fi >> (8 * k)) & 1u); }
        }
    }
}

// backward of the not-fired rows: thread = one 16-byte piece of a row's (dAdir 4 | dY 8) pieces, coalesced
__global__ void __launch_bounds__(256)
step_bwd_unfired_kernel(int64_t n, const uint8_t* __restrict__ flags_s, const float* __restrict__ dAn_s,
                        float* __restrict__ dY_s, float* __restrict__ dAdir_s) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t p = t / 12;
    const int piece = (int)(t - p * 12);
    if (p >= n) return;
    const uint8_t fl_ = flags_s[p];
    if (fl_ & 4) return;                                      // fired rows belong to the tensor-core kernel
    if (piece < 4) {
        const bool living = (fl_ & 1) && (fl_ & 2);
        const float4 g = living ? __ldg(reinterpret_cast<const float4*>(dAn_s + p * NF) + piece) : make_float4(0.f, 0.f, 0.f, 0.f);
        reinterpret_cast<float4*>(dAdir_s + p * NF)[piece] = g;
    } else {
        reinterpret_cast<float4*>(dY_s + p * 2 * NF)[piece - 4] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

}  // namespace sphnca

using namespace sphnca;

extern "C" {

int32_t sphnca_fire_list_workspace_bytes(int64_t N, size_t* bytes) {
    SPHNCA_CHECK_ARG(N >= 0 && bytes, "bad argument");
    *bytes = (size_t)((N + fl::BLOCK_ROWS - 1) / fl::BLOCK_ROWS + 1) * sizeof(int32_t);
    return SPHNCA_OK;
}

static int32_t fire_list_impl(int64_t N, const uint8_t* fire_s, const float* A_s, int32_t use_alpha, int32_t* rows,
                              int32_t* count, float* A_new_s, uint8_t* alive_s, const uint8_t* FLG, uint8_t* flags_s,
                              int32_t fold_static_mask, const uint8_t* carry_sel, void* workspace, size_t workspace_bytes, void* stream_) {
    SPHNCA_CHECK_ARG(N >= 0 && N < (int64_t(1) << 31), "bad N");
    SPHNCA_CHECK_ARG(count, "NULL argument");
    cudaStream_t stream = as_stream(stream_);
    if (N == 0) { SPHNCA_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), stream)); return SPHNCA_OK; }
    SPHNCA_CHECK_ARG(fire_s && A_s && rows, "NULL argument");
    SPHNCA_CHECK_ARG((A_new_s == nullptr) == (alive_s == nullptr), "A_new_s and alive_s: both or neither");
    SPHNCA_CHECK_ARG(((uintptr_t)fire_s & 3) == 0, "fire_s must be 4-byte aligned");
    SPHNCA_CHECK_ARG(!fold_static_mask || (FLG != nullptr && !use_alpha), "the static life mask needs FLG and no alpha channel");
    const int nb = (int)((N + fl::BLOCK_ROWS - 1) / fl::BLOCK_ROWS);
    if (!workspace || workspace_bytes < (size_t)(nb + 1) * sizeof(int32_t)) { set_error("fire-list workspace too small"); return SPHNCA_ERR_WORKSPACE; }
    int32_t* bc = reinterpret_cast<int32_t*>(workspace);
    SPHNCA_LAUNCH(fire_count_kernel, nb, fl::THREADS, 0, stream, N, fire_s, bc);
    SPHNCA_LAUNCH(fire_scan_kernel, 1, 1024, 0, stream, nb, bc, count);
    SPHNCA_LAUNCH(fire_fill_kernel, nb, fl::THREADS, 0, stream, N, fire_s, bc, A_s, use_alpha, rows, A_new_s, alive_s, FLG, flags_s,
                  fold_static_mask ? 1 : 0, carry_sel);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_fire_list(int64_t N, const uint8_t* fire_s, const float* A_s, int32_t use_alpha, int32_t* rows,
                         int32_t* count, float* A_new_s, uint8_t* alive_s, const uint8_t* FLG, uint8_t* flags_s,
                         int32_t fold_static_mask, void* workspace, size_t workspace_bytes, void* stream_) {
    SPHNCA_CHECK_ARG(FLG == nullptr || flags_s != nullptr, "flags_s is required with FLG");
    return fire_list_impl(N, fire_s, A_s, use_alpha, rows, count, A_new_s, alive_s, FLG, flags_s, fold_static_mask, nullptr, workspace,
                          workspace_bytes, stream_);
}

int32_t sphnca_fire_list_draw(const sphnca_grid* g, const float* u, float rate, uint8_t* fire_s, const float* A_s,
                              int32_t use_alpha, int32_t* rows, int32_t* count, float* A_new_s, uint8_t* alive_s,
                              const uint8_t* FLG, uint8_t* flags_s, int32_t fold_static_mask, void* workspace,
                              size_t workspace_bytes, void* stream_) {
    GridView gv; { int rc = make_view(g, &gv, true); if (rc) return rc; }
    const int64_t N = gv.n;
    SPHNCA_CHECK_ARG(count, "NULL argument");
    cudaStream_t stream = as_stream(stream_);
    if (N == 0) { SPHNCA_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), stream)); return SPHNCA_OK; }
    SPHNCA_CHECK_ARG(u && fire_s && A_s && rows && A_new_s && alive_s, "NULL argument");
    SPHNCA_CHECK_ARG(((uintptr_t)fire_s & 3) == 0, "fire_s must be 4-byte aligned");
    SPHNCA_CHECK_ARG(FLG == nullptr || flags_s != nullptr, "flags_s is required with FLG");
    SPHNCA_CHECK_ARG(!fold_static_mask || (FLG != nullptr && !use_alpha), "the static life mask needs FLG and no alpha channel");
    const int nb = (int)((N + fl::BLOCK_ROWS - 1) / fl::BLOCK_ROWS);
    if (!workspace || workspace_bytes < (size_t)(nb + 1) * sizeof(int32_t)) { set_error("fire-list workspace too small"); return SPHNCA_ERR_WORKSPACE; }
    int32_t* bc = reinterpret_cast<int32_t*>(workspace);
    SPHNCA_LAUNCH(fire_draw_count_kernel, nb, fl::THREADS, 0, stream, N, u, rate, gv.sorted_id, fire_s, bc);
    SPHNCA_LAUNCH(fire_scan_kernel, 1, 1024, 0, stream, nb, bc, count);
    SPHNCA_LAUNCH(fire_fill_kernel, nb, fl::THREADS, 0, stream, N, fire_s, bc, A_s, use_alpha, rows, A_new_s, alive_s, FLG, flags_s,
                  fold_static_mask ? 1 : 0, nullptr);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_step_fwd_fired(const sphnca_grid* g, const uint8_t* fire_s, const float* A_s, const int32_t* group_slots,
                              const int64_t* group_base, const int32_t* J, const float* WP, const uint8_t* FLG,
                              const sphnca_step_params* p, const uint8_t* carry_sel, int32_t write_flags, int64_t max_fired,
                              int32_t* rows, int32_t* count, float* Y_c, float* A_c, float* A_new_s, uint8_t* alive_s, float* O_c,
                              uint8_t* flags_s, uint32_t* relu_c, void* workspace, size_t workspace_bytes, void* stream) {
    SPHNCA_CHECK_ARG(g && p && !p->use_alpha, "sphnca_step_fwd_fired is the no-alpha step (static life mask)");
    SPHNCA_CHECK_ARG(A_new_s && alive_s && flags_s && FLG, "NULL argument");
    const int64_t N = g->num_particles;
    // list of the fired rows; the not-fired rows (of carry_sel, when given) are carried over; the static mask bits of ALL rows are
    // written when write_flags is set (the first list of a step)
    int rc = fire_list_impl(N, fire_s, A_s, 0, rows, count, A_new_s, alive_s, FLG, write_flags ? flags_s : nullptr, 1, carry_sel,
                            workspace, workspace_bytes, stream);
    if (rc) return rc;
    if (N == 0) return SPHNCA_OK;
    // the list kernels size their grids for an upper bound of the list length: the caller may know a much smaller one
    sphnca_grid gb = *g;
    if (max_fired > 0 && max_fired < N) gb.num_particles = max_fired;
    rc = sphnca_step_perceive_pairs_fired(&gb, rows, count, group_slots, group_base, J, WP, A_s, p, Y_c, A_c, stream);
    if (rc) return rc;
    return sphnca_step_mlp_list(gb.num_particles, rows, count, A_c, Y_c, p, A_new_s, alive_s, O_c, flags_s, relu_c, nullptr, nullptr, 3, stream);
}

int32_t sphnca_fire_bytes_counter(int64_t N, const int64_t* gid_s, int64_t seed, float rate, const uint8_t* bnd_s, uint8_t* fire_s,
                                  uint8_t* fire_b, uint8_t* fire_i, void* stream_) {
    SPHNCA_CHECK_ARG(N >= 0 && N < (int64_t(1) << 31), "bad N");
    if (N == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(gid_s && fire_s, "NULL argument");
    SPHNCA_CHECK_ARG((fire_b == nullptr) == (fire_i == nullptr) && (fire_b == nullptr || bnd_s != nullptr), "fire_b / fire_i / bnd_s: all or none");
    SPHNCA_CHECK_ARG((((uintptr_t)fire_s | (uintptr_t)fire_b | (uintptr_t)fire_i) & 3) == 0, "fire byte arrays must be 4-byte aligned");
    const uint64_t seed_term = 0x9E3779B97F4A7C15ull * (uint64_t)(seed + 1);
    const int64_t threads = (N + 3) / 4;
    SPHNCA_LAUNCH(fire_bytes_counter_kernel, (unsigned)((threads + 255) / 256), 256, 0, as_stream(stream_), N, gid_s, seed_term, rate, bnd_s,
                  fire_s, fire_b, fire_i);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_step_bwd_unfired(int64_t N, const uint8_t* flags_s, const float* dA_next_s, float* dY_s, float* dAdir_s,
                                void* stream_) {
    SPHNCA_CHECK_ARG(N >= 0 && N < (int64_t(1) << 31), "bad N");
    if (N == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(flags_s && dA_next_s && dY_s && dAdir_s, "NULL argument");
    const int64_t threads = N * 12;
    SPHNCA_LAUNCH(step_bwd_unfired_kernel, (unsigned)((threads + 255) / 256), 256, 0, as_stream(stream_), N, flags_s, dA_next_s, dY_s, dAdir_s);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

}  // extern "C"
