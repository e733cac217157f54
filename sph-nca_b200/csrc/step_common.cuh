// Definitions shared by the CUDA-core (step.cu) and tcgen05 (step_tc.cu) implementations of the fused NCA step.
#pragma once
#include "common.cuh"

namespace sphnca {

constexpr int NF = 16;        // channels
constexpr int NIN = 48;       // MLP input  [A | dA/dx | dA/dy]
constexpr int NH = 256;       // hidden width
constexpr int NO_PAD = 36;    // MLP output 33 (gated) / 16 (orig), padded for float4 rows
constexpr int kStepThreads = 256;

struct StepConsts {
    float h, h2;
    float cw;          // poly6 normalisation
    float kg;          // spiky normalisation * perception scale
    int use_alpha, rule, n_out;
    float orig_rate;
    int dy_ilv;        // sphnca_step_params.dy_interleaved
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <int D>
__device__ __forceinline__ float pair_d2(const float4& a, const float4& b, float& rx, float& ry) {
    rx = b.x - a.x; ry = b.y - a.y;
    float d2 = fmaf(ry, ry, rx * rx);
    if (D == 3) { const float rz = b.z - a.z; d2 = fmaf(rz, rz, d2); }
    return d2;
}

__device__ __forceinline__ void load16(const float* __restrict__ p, float (&o)[NF]) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int k = 0; k < 4; ++k) { float4 t = __ldg(p4 + k); o[4*k] = t.x; o[4*k+1] = t.y; o[4*k+2] = t.z; o[4*k+3] = t.w; }
}

inline int make_consts(const sphnca_grid* g, const sphnca_step_params* p, StepConsts* c) {
    if (!p || !p->W1 || !p->b1 || !p->W2 || !p->b2) { set_error("step params / weights are NULL"); return SPHNCA_ERR_INVALID; }
    if (p->update_rule != 0 && p->update_rule != 1) { set_error("update_rule must be 0 (gated) or 1 (orig)"); return SPHNCA_ERR_INVALID; }
    if ((((uintptr_t)p->W1 | (uintptr_t)p->W2) & 15) != 0) { set_error("W1 and W2 must be 16-byte aligned"); return SPHNCA_ERR_INVALID; }
    c->h = (float)g->h; c->h2 = (float)(g->h * g->h);
    c->cw = (float)poly6_norm(g->dim, g->h);
    c->kg = (float)(spiky_norm(g->dim, g->h) * (double)p->perception_scale);
    c->use_alpha = p->use_alpha; c->rule = p->update_rule; c->n_out = p->update_rule == 0 ? 2 * NF + 1 : NF;
    c->orig_rate = p->orig_rate;
    c->dy_ilv = p->dy_interleaved ? 1 : 0;
    return SPHNCA_OK;
}


// which implementation the step uses: 0 = CUDA-core MLP (step.cu), 1 = tcgen05, split forward
// (step_perceive_kernel + step_mlp_tc.cu; default), 2 = tcgen05, fused forward (step_tc.cu)
int mlp_path();
int launch_step_fwd_tc(const GridView& gv, const StepConsts& c, const sphnca_step_params* p, const float* xv4,
                       const float* A_s, const uint8_t* fire_s, float* A_new_s, uint8_t* alive_s, float* Y_s,
                       float* O_s, uint8_t* flags_s, uint32_t* relu_s, cudaStream_t stream);

// rows / count (fire-list mode, firelist.cu): NULL = all N rows in order; else the kernels run over rows[0..*count) and
// N is only the upper bound the grid is sized for
int launch_step_mlp_tc(int64_t N, const StepConsts& c, const sphnca_step_params* p, const float* A_s, const float* Y_s,
                       const uint8_t* fire_s, float* A_new_s, uint8_t* alive_s, float* O_s, uint8_t* flags_s,
                       uint32_t* relu_s, cudaStream_t stream, const int32_t* rows = nullptr, const int32_t* count = nullptr,
                       float* A_c = nullptr, float* Y_c = nullptr, int in_by_list = 0, int fold_mask = 0);
int launch_step_bwd_tc(int64_t N, const StepConsts& c, const sphnca_step_params* p, const float* A_s, const float* O_s,
                       const uint8_t* flags_s, const uint32_t* relu_s, const float* dA_next_s, float* dO_s, float* dY_s,
                       float* dAdir_s, cudaStream_t stream, const int32_t* rows = nullptr, const int32_t* count = nullptr);
int wgrad_tc_streams(int64_t N);
int launch_mlp_wgrad_tc(int64_t N, const float* A_s, const float* Y_s, const float* dO_s, const sphnca_step_params* p,
                        int n_out, float* partial, int part_floats, int streams, cudaStream_t stream,
                        const int32_t* count = nullptr, int acc_in = 0);

}  // namespace sphnca
