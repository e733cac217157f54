// Fire lists: the MLP half of an NCA step only matters for the particles whose fire-mask draw came up.
//
// Reference code/nca.py:100-107 evaluates the MLP for EVERY particle and then keeps the old state where the
// stochastic update mask is off (`torch.where(update_mask, nA, A)`): for those particles the MLP output is
// discarded in the forward, and in the backward no gradient reaches the MLP (dL/dO == 0, so they add exactly 0
// to dW1, db1, dW2, db2 and to dL/dY).  With the default fire rate 0.5 that is half of all tensor-core work.
//
//   fire_list        rows[0..count) = sorted positions of the fired particles, in increasing order (ordered
//                    compaction in ONE launch: block counts chained by a look-back over the preceding blocks; deterministic);
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

// ---- one launch: fire bytes (drawn or given) -> ordered list + carry-over of the not-fired rows ------------------------------
// Single-pass ordered compaction: a block takes a ticket (blocks are numbered in the order they START, so every predecessor is
// already running -- no deadlock whatever the residency), counts its fired rows, publishes the count in its state word and sums
// its predecessors' words (see below).  Sums of integers: the list is the same whatever the timing.  Replaces three launches (draw + block counts | one-block scan |
// fill); the not-fired rows' 64-byte state copies are done as coalesced 16-byte pieces (4 lanes per row) instead of a thread
// per row.  State words and the two counters live in the caller's workspace: ZERO before the first use, left zero by every
// launch (the block that finishes last clears them).
namespace fl {
constexpr unsigned long long FLAG_AGG = 1ull << 32, VALUE_MASK = 0xffffffffull;
__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
}  // namespace fl

template <bool DRAW>
__global__ void __launch_bounds__(fl::THREADS)
fire_list_onepass_kernel(int64_t n, int nb, const float* __restrict__ u, float rate, const int32_t* __restrict__ sorted_id,
                         uint8_t* __restrict__ fire_s, const float* __restrict__ A_s, int use_alpha, int32_t* __restrict__ rows,
                         int32_t* __restrict__ count, float* __restrict__ A_new_s, uint8_t* __restrict__ alive_s,
                         const uint8_t* __restrict__ FLG, uint8_t* __restrict__ flags_s, int fold_mask,
                         const uint8_t* __restrict__ carry_sel, unsigned long long* __restrict__ state, int* __restrict__ ctr) {
    using namespace fl;
    __shared__ int s_ticket, s_excl, wsum[THREADS / 32];
    __shared__ uint8_t s_copy[BLOCK_ROWS];                    // per row: 0 nothing to copy, 1 copy, 2 copy times zero (dead, static mask folded)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_ticket = atomicAdd(&ctr[0], 1);
    __syncthreads();
    const int b = s_ticket;
    const int64_t base = (int64_t)b * BLOCK_ROWS + (int64_t)tid * PER_THREAD;
    // ---- fire bits of this thread's 4 rows
    uint32_t packed = 0;
    if (DRAW) {
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k) {
            const int64_t p = base + k;
            packed |= ((p < n && __ldg(u + __ldg(sorted_id + p)) <= rate) ? 1u : 0u) << (8 * k);
        }
        if (base + PER_THREAD <= n) *reinterpret_cast<uint32_t*>(fire_s + base) = packed;
        else for (int k = 0; k < PER_THREAD; ++k) if (base + k < n) fire_s[base + k] = (uint8_t)((packed >> (8 * k)) & 1u);
    } else {
        if (base + PER_THREAD <= n) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(fire_s + base);
            packed = ((w & 0xffu) ? 1u : 0u) | ((w & 0xff00u) ? 0x100u : 0u) | ((w & 0xff0000u) ? 0x10000u : 0u) | ((w & 0xff000000u) ? 0x1000000u : 0u);
        } else {
            for (int k = 0; k < PER_THREAD; ++k) if (base + k < n && fire_s[base + k] != 0) packed |= 1u << (8 * k);
        }
    }
    const int c = __popc(packed);
    // ---- ranks inside the block
    int x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    int before = x - c, total = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) { const int t = wsum[w]; if (w < warp) before += t; total += t; }
    // ---- the block's count goes into its state word at once; the sums over the preceding blocks are only read after the carry-over
    // copies below (nothing but the list positions depends on them), when every block that started earlier has published its own
    if (tid == 0) fl::st_state(state + b, fl::FLAG_AGG | (unsigned long long)(uint32_t)total);
    // ---- per-row work that needs no prefix: static-mask flag bytes, what to copy
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) {
        const int64_t p = base + k;
        uint8_t cp = 0;
        if (p < n) {
            // static life mask (no alpha channel): the previous-mask bit of every row; fold_mask: the new mask is the same bit
            // (step_mask_static_kernel), so bit 1 is set here too and dead rows are zeroed below / by the MLP kernel
            const bool live = FLG == nullptr || FLG[p] != 0;
            if (FLG != nullptr && flags_s != nullptr) flags_s[p] = live ? (fold_mask ? 3 : 1) : 0;
            const bool fired = (packed >> (8 * k)) & 1u;
            // not fired: the state is carried over unchanged (reference nca.py:107); fire bit of the flags stays 0.
            // carry_sel: rows carried over by the other list of the step are skipped (multigpu.py)
            if (!fired && A_new_s != nullptr && (carry_sel == nullptr || carry_sel[p] != 0)) cp = (fold_mask && !live) ? 2 : 1;
        }
        s_copy[tid * PER_THREAD + k] = cp;
    }
    __syncthreads();
    // ---- carry-over copies: 4 lanes per row, 16 bytes each
    if (A_new_s != nullptr) {
        const float4* a4 = reinterpret_cast<const float4*>(A_s) + (int64_t)b * BLOCK_ROWS * 4;
        float4* o4 = reinterpret_cast<float4*>(A_new_s) + (int64_t)b * BLOCK_ROWS * 4;
        uint8_t* al = alive_s + (int64_t)b * BLOCK_ROWS;
#pragma unroll 8
        for (int i = tid; i < BLOCK_ROWS * 4; i += THREADS) {
            const int r = i >> 2;
            const uint8_t cp = s_copy[r];
            if (cp == 0) continue;
            float4 v = __ldg(a4 + i);
            if ((i & 3) == 0) al[r] = use_alpha ? (v.w > 0.1f ? 1 : 0) : 1;
            if (cp == 2) v = make_float4(v.x * 0.f, v.y * 0.f, v.z * 0.f, v.w * 0.f);   // the product with 0 of step_mask_static_kernel, bit for bit
            o4[i] = v;
        }
    }
    // ---- exclusive prefix of the block: the sum over the preceding blocks' state words is read
    // back in PARALLEL in two levels (<= 255 words of its own group of 256 blocks + one total per preceding group, published by
    // each group's last block): two global-memory round trips however many blocks there are.  (The textbook chained look-back
    // serialises here: all blocks of a launch run at once, none has a prefix yet, and block b would walk back b/32 windows --
    // 12 us for 512 blocks.)
    {
        unsigned long long* gstate = state + nb;
        const int g = b >> 8, b0 = g << 8;
        auto spin_value = [&](const unsigned long long* w) -> int {
            unsigned long long v = ld_state(w);
            for (uint32_t spins = 0; (v >> 32) == 0; v = ld_state(w))
                if (++spins > (1u << 26)) asm volatile("trap;");                        // a protocol bug traps instead of hanging the GPU
            return (int)(v & VALUE_MASK);
        };
        auto block_sum = [&](int v) -> int {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            __syncthreads();                                                            // wsum free again
            if (lane == 0) wsum[warp] = v;
            __syncthreads();
            int t = 0;
#pragma unroll
            for (int w = 0; w < THREADS / 32; ++w) t += wsum[w];
            return t;
        };
        const int in_group = block_sum(tid < b - b0 ? spin_value(state + b0 + tid) : 0);
        if ((b & 255) == 255 && tid == 0) st_state(gstate + g, FLAG_AGG | (unsigned long long)(uint32_t)(in_group + total));
        int prev = 0;
        for (int gi = tid; gi < g; gi += THREADS) prev += spin_value(gstate + gi);
        const int excl = in_group + block_sum(prev);
        if (tid == 0) { s_excl = excl; if (b == nb - 1) count[0] = excl + total; }
    }
    __syncthreads();
    // ---- the list
    {
        int slot = s_excl + before;
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k) if ((packed >> (8 * k)) & 1u) rows[slot++] = (int32_t)(base + k);
    }
    // ---- the block that finishes last leaves the workspace zeroed for the next launch
    __syncthreads();
    if (tid == 0) { __threadfence(); s_ticket = atomicAdd(&ctr[1], 1); }
    __syncthreads();
    if (s_ticket == nb - 1) {
        for (int i = tid; i < nb + ((nb + 255) >> 8); i += THREADS) st_state(state + i, 0ull);
        if (tid == 0) { ctr[0] = 0; ctr[1] = 0; }
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

// state words of the blocks, of the groups of 256 blocks, {ticket, finished} counters
static size_t fl_workspace_bytes(int64_t N) {
    const size_t nb = (size_t)((N + fl::BLOCK_ROWS - 1) / fl::BLOCK_ROWS);
    return (nb + ((nb + 255) >> 8) + 1) * sizeof(unsigned long long);
}

extern "C" {

int32_t sphnca_fire_list_workspace_bytes(int64_t N, size_t* bytes) {
    SPHNCA_CHECK_ARG(N >= 0 && bytes, "bad argument");
    *bytes = fl_workspace_bytes(N);
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
    if (!workspace || workspace_bytes < fl_workspace_bytes(N)) { set_error("fire-list workspace too small"); return SPHNCA_ERR_WORKSPACE; }
    SPHNCA_CHECK_ARG(((uintptr_t)workspace & 7) == 0, "the fire-list workspace must be 8-byte aligned");
    unsigned long long* state = reinterpret_cast<unsigned long long*>(workspace);
    SPHNCA_LAUNCH(fire_list_onepass_kernel<false>, nb, fl::THREADS, 0, stream, N, nb, (const float*)nullptr, 0.f, (const int32_t*)nullptr,
                  const_cast<uint8_t*>(fire_s), A_s, use_alpha, rows, count, A_new_s, alive_s, FLG, flags_s, fold_static_mask ? 1 : 0, carry_sel,
                  state, reinterpret_cast<int*>(state + nb + ((nb + 255) >> 8)));
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
    if (!workspace || workspace_bytes < fl_workspace_bytes(N)) { set_error("fire-list workspace too small"); return SPHNCA_ERR_WORKSPACE; }
    SPHNCA_CHECK_ARG(((uintptr_t)workspace & 7) == 0, "the fire-list workspace must be 8-byte aligned");
    unsigned long long* state = reinterpret_cast<unsigned long long*>(workspace);
    SPHNCA_LAUNCH(fire_list_onepass_kernel<true>, nb, fl::THREADS, 0, stream, N, nb, u, rate, gv.sorted_id, fire_s, A_s, use_alpha, rows, count,
                  A_new_s, alive_s, FLG, flags_s, fold_static_mask ? 1 : 0, (const uint8_t*)nullptr, state, reinterpret_cast<int*>(state + nb + ((nb + 255) >> 8)));
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
