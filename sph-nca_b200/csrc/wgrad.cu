// Weight gradients of the 48->256->n_out MLP, summed over particles:
//   dW1[n][k] += sum_p dH[p][n] * y[p][k]      db1[n] += sum_p dH[p][n]
//   dW2[o][n] += sum_p dO[p][o] * H[p][n]      db2[o] += sum_p dO[p][o]
// Replaces torch autograd's two skinny cuBLAS GEMMs + reductions.  H and dH arrive TRANSPOSED and
// tile-blocked ([N/32][256][32], written coalesced by step_bwd_kernel; one 32-particle tile is a
// contiguous 32 KB block), so thread n streams its own 128-byte row piece straight into registers
// (LDG.128) and keeps its 48 + 36 + 1 outputs in registers; the per-particle perception
// row y and dO row are broadcast from a double-buffered smem tile.  Two-stage, fixed-order
// reduction (per-CTA partials, then a tree over CTAs): deterministic, no float atomics.
#include "step_common.cuh"

namespace sphnca {

constexpr int WG_THREADS = NH;      // one thread per hidden unit
constexpr int WG_TILE = 32;         // particles per smem tile
constexpr int WG_PART = NH * NIN + NH + NO_PAD * NH + NO_PAD;   // floats per partial: dW1 | db1 | dW2 | db2

struct WgTile {
    float y[WG_TILE][NIN];          // [A | dA/dx | dA/dy]
    float dO[WG_TILE][NO_PAD];
};

__device__ __forceinline__ void wg_load_tile(WgTile& t, const float* __restrict__ A_s, const float* __restrict__ Y_s,
                                             const float* __restrict__ dO_s, int64_t p0, int64_t p_end) {
    // 32 particles x (4 + 8 + 9) float4 = 672 float4, 256 threads
    for (int i = threadIdx.x; i < WG_TILE * 21; i += WG_THREADS) {
        const int j = i / 21, q = i - j * 21;
        const int64_t p = p0 + j;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < p_end) {
            if (q < 4) v = __ldg(reinterpret_cast<const float4*>(A_s + p * NF) + q);
            else if (q < 12) v = __ldg(reinterpret_cast<const float4*>(Y_s + p * 2 * NF) + (q - 4));
            else v = __ldg(reinterpret_cast<const float4*>(dO_s + p * NO_PAD) + (q - 12));
        }
        if (q < 12) *reinterpret_cast<float4*>(&t.y[j][4 * q]) = v;
        else *reinterpret_cast<float4*>(&t.dO[j][4 * (q - 12)]) = v;
    }
}

__global__ void __launch_bounds__(WG_THREADS, 2)
mlp_wgrad_kernel(int64_t n_total, const float* __restrict__ A_s, const float* __restrict__ Y_s,
                 const float* __restrict__ dO_s, const float* __restrict__ Ht, const float* __restrict__ dHt,
                 float* __restrict__ partial) {
    __shared__ __align__(16) WgTile tiles[2];
    const int n = threadIdx.x;
    // contiguous particle range of this CTA, in whole tiles
    const int64_t tiles_total = (n_total + WG_TILE - 1) / WG_TILE;
    const int64_t t0 = tiles_total * blockIdx.x / gridDim.x, t1 = tiles_total * (blockIdx.x + 1) / gridDim.x;
    float aW1[NIN], aW2[NO_PAD], ab1 = 0.f, ab2 = 0.f;
#pragma unroll
    for (int k = 0; k < NIN; ++k) aW1[k] = 0.f;
#pragma unroll
    for (int o = 0; o < NO_PAD; ++o) aW2[o] = 0.f;
    if (t0 < t1) wg_load_tile(tiles[0], A_s, Y_s, dO_s, t0 * WG_TILE, n_total);
    __syncthreads();
    for (int64_t t = t0; t < t1; ++t) {
        const int buf = (int)((t - t0) & 1);
        if (t + 1 < t1) wg_load_tile(tiles[buf ^ 1], A_s, Y_s, dO_s, (t + 1) * WG_TILE, n_total);
        const WgTile& tl = tiles[buf];
        const int64_t p0 = t * WG_TILE;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float h[16], dh[16];
            // tile t is a contiguous [256][32] block; the scratch is allocated in whole tiles and the
            // tail entries (p >= n_total) multiply zeroed y / dO rows, but may hold non-finite garbage
            const size_t ro = ((size_t)t * NH + n) * 32 + half * 16;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(Ht + ro) + q);
                const float4 b = __ldg(reinterpret_cast<const float4*>(dHt + ro) + q);
                h[4 * q] = a.x; h[4 * q + 1] = a.y; h[4 * q + 2] = a.z; h[4 * q + 3] = a.w;
                dh[4 * q] = b.x; dh[4 * q + 1] = b.y; dh[4 * q + 2] = b.z; dh[4 * q + 3] = b.w;
            }
            if (p0 + half * 16 + 16 > n_total) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (p0 + half * 16 + j >= n_total) { h[j] = 0.f; dh[j] = 0.f; }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4* y4 = reinterpret_cast<const float4*>(&tl.y[half * 16 + j][0]);
                const float4* o4 = reinterpret_cast<const float4*>(&tl.dO[half * 16 + j][0]);
                ab1 += dh[j];
#pragma unroll
                for (int k = 0; k < NIN / 4; ++k) {
                    const float4 v = y4[k];
                    aW1[4 * k] = fmaf(dh[j], v.x, aW1[4 * k]); aW1[4 * k + 1] = fmaf(dh[j], v.y, aW1[4 * k + 1]);
                    aW1[4 * k + 2] = fmaf(dh[j], v.z, aW1[4 * k + 2]); aW1[4 * k + 3] = fmaf(dh[j], v.w, aW1[4 * k + 3]);
                }
#pragma unroll
                for (int k = 0; k < NO_PAD / 4; ++k) {
                    const float4 v = o4[k];
                    aW2[4 * k] = fmaf(h[j], v.x, aW2[4 * k]); aW2[4 * k + 1] = fmaf(h[j], v.y, aW2[4 * k + 1]);
                    aW2[4 * k + 2] = fmaf(h[j], v.z, aW2[4 * k + 2]); aW2[4 * k + 3] = fmaf(h[j], v.w, aW2[4 * k + 3]);
                }
            }
        }
        if (n < NO_PAD) {
#pragma unroll 8
            for (int j = 0; j < WG_TILE; ++j) ab2 += tl.dO[j][n];
        }
        __syncthreads();
    }
    float* out = partial + (size_t)blockIdx.x * WG_PART;
#pragma unroll
    for (int k = 0; k < NIN; ++k) out[n * NIN + k] = aW1[k];
    out[NH * NIN + n] = ab1;
#pragma unroll
    for (int o = 0; o < NO_PAD; ++o) out[NH * NIN + NH + o * NH + n] = aW2[o];
    if (n < NO_PAD) out[NH * NIN + NH + NO_PAD * NH + n] = ab2;
}

// stage 2: fixed-order sum over the CTA partials, accumulated into the gradient tensors
__global__ void __launch_bounds__(256)
mlp_wgrad_reduce_kernel(const float* __restrict__ partial, int num_partials, int n_out, float* __restrict__ dW1,
                        float* __restrict__ db1, float* __restrict__ dW2, float* __restrict__ db2, int w1_transposed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= WG_PART) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int g = 0;
    for (; g + 4 <= num_partials; g += 4) {
        s0 += partial[(size_t)g * WG_PART + i]; s1 += partial[(size_t)(g + 1) * WG_PART + i];
        s2 += partial[(size_t)(g + 2) * WG_PART + i]; s3 += partial[(size_t)(g + 3) * WG_PART + i];
    }
    for (; g < num_partials; ++g) s0 += partial[(size_t)g * WG_PART + i];
    const float s = (s0 + s1) + (s2 + s3);
    if (i < NH * NIN) {
        if (w1_transposed) dW1[(i % NH) * NIN + i / NH] += s;      // tensor-core partials hold dW1 as [k][h]
        else dW1[i] += s;
    }
    else if (i < NH * NIN + NH) db1[i - NH * NIN] += s;
    else if (i < NH * NIN + NH + NO_PAD * NH) {
        const int r = i - NH * NIN - NH, o = r / NH, n = r - o * NH;
        if (o < n_out) dW2[o * NH + n] += s;
    } else {
        const int o = i - NH * NIN - NH - NO_PAD * NH;
        if (o < n_out) db2[o] += s;
    }
}

static int wg_grid(int64_t N) {
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) num_sms = 148;
    }
    int64_t tiles = (N + WG_TILE - 1) / WG_TILE;
    int64_t g = 2 * (int64_t)num_sms;
    if (g > tiles) g = tiles;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace sphnca

using namespace sphnca;

extern "C" {

int32_t sphnca_mlp_wgrad_workspace_bytes(int64_t N, size_t* bytes) {
    SPHNCA_CHECK_ARG(N >= 0 && bytes, "bad argument");
    const int a = wg_grid(N), b = wgrad_tc_streams(N);
    *bytes = (size_t)(a > b ? a : b) * WG_PART * sizeof(float);
    return SPHNCA_OK;
}

int32_t sphnca_mlp_wgrad(int64_t N, const float* A_s, const float* Y_s, const float* dO_s, const float* Ht,
                         const float* dHt, const sphnca_step_params* p, float* dW1, float* db1, float* dW2, float* db2,
                         void* workspace, size_t workspace_bytes, void* stream_) {
    SPHNCA_CHECK_ARG(N >= 0 && p && (p->update_rule == 0 || p->update_rule == 1), "bad N / params");
    const int n_out = p->update_rule == 0 ? 2 * NF + 1 : NF;
    if (N == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(A_s && Y_s && dO_s && dW1 && db1 && dW2 && db2, "NULL argument");
    SPHNCA_CHECK_ARG((Ht == nullptr) == (dHt == nullptr), "Ht and dHt must be given together");
    if (Ht == nullptr) {
        // tensor-core path: hidden layer recomputed in the transposed orientation (wgrad_tc.cu)
        SPHNCA_CHECK_ARG(p->W1 && p->b1 && p->W2, "weights are required by the recompute path");
        SPHNCA_CHECK_ARG(((uintptr_t)p->W1 & 15) == 0, "W1 must be 16-byte aligned");
        const int streams = wgrad_tc_streams(N);
        if (!workspace || workspace_bytes < (size_t)streams * WG_PART * sizeof(float)) { set_error("wgrad workspace too small"); return SPHNCA_ERR_WORKSPACE; }
        float* partial = reinterpret_cast<float*>(workspace);
        int rc = launch_mlp_wgrad_tc(N, A_s, Y_s, dO_s, p, n_out, partial, WG_PART, streams, as_stream(stream_));
        if (rc) return rc;
        SPHNCA_LAUNCH(mlp_wgrad_reduce_kernel, (WG_PART + 255) / 256, 256, 0, as_stream(stream_), partial, streams, n_out, dW1, db1, dW2, db2, 1);
        SPHNCA_LAUNCH_CHECK();
        return SPHNCA_OK;
    }
    const int grid = wg_grid(N);
    if (!workspace || workspace_bytes < (size_t)grid * WG_PART * sizeof(float)) { set_error("wgrad workspace too small"); return SPHNCA_ERR_WORKSPACE; }
    cudaStream_t stream = as_stream(stream_);
    float* partial = reinterpret_cast<float*>(workspace);
    SPHNCA_LAUNCH(mlp_wgrad_kernel, grid, WG_THREADS, 0, stream, N, A_s, Y_s, dO_s, Ht, dHt, partial);
    SPHNCA_LAUNCH(mlp_wgrad_reduce_kernel, (WG_PART + 255) / 256, 256, 0, stream, partial, grid, n_out, dW1, db1, dW2, db2, 0);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

int32_t sphnca_mlp_wgrad_list(int64_t N, const int32_t* count, const float* A_c, const float* Y_c, const float* dO_c,
                              const sphnca_step_params* p, float* dW1, float* db1, float* dW2, float* db2,
                              void* workspace, size_t workspace_bytes, void* stream_) {
    SPHNCA_CHECK_ARG(N >= 0 && p && (p->update_rule == 0 || p->update_rule == 1), "bad N / params");
    const int n_out = p->update_rule == 0 ? 2 * NF + 1 : NF;
    if (N == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(count && A_c && Y_c && dO_c && dW1 && db1 && dW2 && db2, "NULL argument");
    SPHNCA_CHECK_ARG(p->W1 && p->b1 && p->W2, "weights are required by the recompute path");
    SPHNCA_CHECK_ARG(((uintptr_t)p->W1 & 15) == 0, "W1 must be 16-byte aligned");
    const int streams = wgrad_tc_streams(N);
    if (!workspace || workspace_bytes < (size_t)streams * WG_PART * sizeof(float)) { set_error("wgrad workspace too small"); return SPHNCA_ERR_WORKSPACE; }
    float* partial = reinterpret_cast<float*>(workspace);
    int rc = launch_mlp_wgrad_tc(N, A_c, Y_c, dO_c, p, n_out, partial, WG_PART, streams, as_stream(stream_), count);
    if (rc) return rc;
    SPHNCA_LAUNCH(mlp_wgrad_reduce_kernel, (WG_PART + 255) / 256, 256, 0, as_stream(stream_), partial, streams, n_out, dW1, db1, dW2, db2, 1);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

// The same in two halves, for a rollout's backward: every step ADDS to the per-stream partials (accumulate = 0 on the first
// call of a series: plain stores), and ONE fixed-order reduction at the end adds them to the parameter gradients -- instead of a
// 6 MB reduction pass per step.  All calls of a series use the same N (it fixes the number of streams) and workspace.
int32_t sphnca_mlp_wgrad_list_partial(int64_t N, const int32_t* count, const float* A_c, const float* Y_c, const float* dO_c,
                                      const sphnca_step_params* p, int32_t accumulate, void* workspace, size_t workspace_bytes,
                                      void* stream_) {
    SPHNCA_CHECK_ARG(N >= 0 && p && (p->update_rule == 0 || p->update_rule == 1), "bad N / params");
    const int n_out = p->update_rule == 0 ? 2 * NF + 1 : NF;
    if (N == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(count && A_c && Y_c && dO_c, "NULL argument");
    SPHNCA_CHECK_ARG(p->W1 && p->b1 && p->W2, "weights are required by the recompute path");
    SPHNCA_CHECK_ARG(((uintptr_t)p->W1 & 15) == 0, "W1 must be 16-byte aligned");
    const int streams = wgrad_tc_streams(N);
    if (!workspace || workspace_bytes < (size_t)streams * WG_PART * sizeof(float)) { set_error("wgrad workspace too small"); return SPHNCA_ERR_WORKSPACE; }
    return launch_mlp_wgrad_tc(N, A_c, Y_c, dO_c, p, n_out, reinterpret_cast<float*>(workspace), WG_PART, streams, as_stream(stream_), count,
                               accumulate ? 1 : 0);
}

int32_t sphnca_mlp_wgrad_reduce(int64_t N, const sphnca_step_params* p, float* dW1, float* db1, float* dW2, float* db2,
                                const void* workspace, size_t workspace_bytes, void* stream_) {
    SPHNCA_CHECK_ARG(N >= 0 && p && (p->update_rule == 0 || p->update_rule == 1), "bad N / params");
    const int n_out = p->update_rule == 0 ? 2 * NF + 1 : NF;
    if (N == 0) return SPHNCA_OK;
    SPHNCA_CHECK_ARG(dW1 && db1 && dW2 && db2, "NULL argument");
    const int streams = wgrad_tc_streams(N);
    if (!workspace || workspace_bytes < (size_t)streams * WG_PART * sizeof(float)) { set_error("wgrad workspace too small"); return SPHNCA_ERR_WORKSPACE; }
    SPHNCA_LAUNCH(mlp_wgrad_reduce_kernel, (WG_PART + 255) / 256, 256, 0, as_stream(stream_), reinterpret_cast<const float*>(workspace), streams,
                  n_out, dW1, db1, dW2, db2, 1);
    SPHNCA_LAUNCH_CHECK();
    return SPHNCA_OK;
}

}  // extern "C"
