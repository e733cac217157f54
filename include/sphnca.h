/*
 * sphnca.h -- C ABI of libsphnca.so: the B200 (sm_100a) implementation of SPH-NCA's per-timestep
 * particle update.  This is the drop-in boundary: every entry point below replaces
 * one Numba-CUDA kernel launch (plus its host-side bookkeeping) of the reference's `sphops`
 * package, and is what a binding in the reference (`code/sphops/ops/func.py`, ctypes) would call.
 *
 * Conventions
 *   - plain C, no torch types; all array arguments are DEVICE pointers unless marked "host";
 *   - the caller owns every buffer (outputs are caller-allocated, nothing is allocated inside);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*); no device syncs;
 *   - return value 0 = success, <0 = error (message via sphnca_last_error(), thread local);
 *   - fp32 particle data, like every shipped configuration of the reference
 *     (`func.py:12-16` allows fp16/fp64 but nothing uses them);
 *   - "original order" = the caller's packed particle order (reference `packed_x`);
 *     "sorted order"   = cell-sorted order: sorted position p holds original particle
 *                        sorted_id[p]; sections stay contiguous and in order.
 *
 * Data layout in HBM: the neighbour sweeps read particle state from physically
 * cell-sorted arrays (`xv4`: float4 {x,y,z,v} per particle, feature rows contiguous per particle),
 * so the 3^D neighbouring cells of a run of particles are a handful of contiguous ranges.
 */
#ifndef SPHNCA_H
#define SPHNCA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPHNCA_OK 0
#define SPHNCA_ERR_INVALID -1   /* bad argument */
#define SPHNCA_ERR_CUDA -2      /* a CUDA runtime call / launch failed */
#define SPHNCA_ERR_WORKSPACE -3 /* workspace too small */

/* Cell tables of one packed batch, as consumed by every sweep.  Plain struct, passed by pointer.
 * Replaces the reference's `Hashgrid` dataclass on the device side
 * (code/sphops/hashgrid/single_impl.py:127-171). */
typedef struct sphnca_grid {
    int32_t dim;               /* D: 2 or 3 (reference kernels are specialised on DIM) */
    int32_t num_sections;      /* B = len(sections) */
    int64_t num_particles;     /* N = sum(sections) */
    int64_t cells_per_section; /* total_cells = prod(dims) */
    int64_t dims[3];           /* cells per axis, x fastest (single_impl.py:34-44) */
    double h;                  /* cell size == support radius */
    const int32_t* cell_off;   /* [B*cells + 1] global sorted offset of every cell (exclusive scan) */
    const int32_t* sorted_id;  /* [N] original (global packed) index of sorted position p */
    const int32_t* sorted_cell;/* [N] global cell id (section*cells + hash) of sorted position p */
    double period[3];          /* > 0: the cloud is PERIODIC along that axis with this period (== dims * h, so that the hash grid's
                                * modulo wrap, single_impl.py:54-55, is the period): a neighbour found through the wrap is taken at
                                * x_j +- period -- the ghost particles of reference nca.py:200-227 without ghost particles.
                                * Honoured by sphnca_volume_fwd, sphnca_pairs_count / _fill (and, through the lists, by every
                                * list-based step kernel); every other sweep entry rejects a periodic grid.  0 = open (default). */
    /* Ghost-volume emulation for periodic planes (x and y axes; NULL = plain periodic volumes).  The reference builds grid and
     * volumes ONCE on the ghost-padded cloud (nca.py:262-263), so a ghost copy carries the volume of ITS OWN, outward-truncated
     * neighbourhood, not its source's, and own particles near an edge see those volumes.  image_volume [N][9] (cell-sorted;
     * entry (sx+1)*3 + (sy+1) = volume of the copy shifted by (sx, sy) periods, entry 4 = the particle's own volume) is written
     * by sphnca_volume_images and read by sphnca_pairs_fill; a copy (sx, sy) of a particle exists iff, per axis, s = 0, or
     * s = +1 and x <= band_lo, or s = -1 and x >= band_hi (nca.py:182-197: within h of the lower / upper edge). */
    const float* image_volume;
    double band_lo[3], band_hi[3];
} sphnca_grid;

/* 3: fire-list workspace = look-back state (zero before the first use, larger than in 2), sphnca_mlp_wgrad_list_partial /
 * sphnca_mlp_wgrad_reduce added, W2 must be 16-byte aligned, sphnca_step_params.dy_interleaved added (struct grows by 8 bytes).  2: periodic grids, fire lists, pair lists. */
int32_t sphnca_abi_version(void);
const char* sphnca_last_error(void);
/* number of kernels this library has launched so far in this process (all threads) */
int64_t sphnca_launch_count(void);

/* ---- (1) hash-grid build: replaces sphops.initialize_hashgrid (code/sphops/preprocess.py:11-43),
 *      i.e. grid_hash_kernel (single_impl.py:62-91) + torch.sort (preprocess.py:36) +
 *      cell_index_init_kernel (single_impl.py:93-124) for ALL sections in one pass.
 *      Outputs idx/cell_start/cell_end/num_points are the reference's Hashgrid fields, bit-exact
 *      (idx ascending inside a cell == stable sort).  cell_off/sorted_id/sorted_cell feed sphnca_grid. */
int32_t sphnca_hashgrid_workspace_bytes(int64_t N, int32_t B, int64_t cells_per_section, size_t* bytes);
int32_t sphnca_hashgrid_build(const float* x, int64_t N, int32_t D,
                              const int64_t* section_sizes_host, int32_t B, double h,
                              const int64_t* dims_host,
                              int64_t* idx, int32_t* cell_start, int32_t* cell_end, int32_t* num_points,
                              int32_t* cell_off, int32_t* sorted_id, int32_t* sorted_cell,
                              void* workspace, size_t workspace_bytes, void* stream);

/* ---- layout helpers (original <-> sorted order) */
/* xv4[p] = {x[sorted_id[p]], v ? v[sorted_id[p]] : 0}; x is [N,D] fp32, z = 0 when D == 2 */
int32_t sphnca_pack_xv(const sphnca_grid* g, const float* x, const float* v, float* xv4, void* stream);
/* dst[p, :] = src[sorted_id[p], :], rows of `row` floats */
int32_t sphnca_pack_rows(const sphnca_grid* g, const float* src, int32_t row, float* dst, void* stream);
/* dst[sorted_id[p], :] = src[p, :] */
int32_t sphnca_unpack_rows(const sphnca_grid* g, const float* src, int32_t row, float* dst, void* stream);
/* same for 1-byte rows (fire masks) */
int32_t sphnca_pack_bytes(const sphnca_grid* g, const uint8_t* src, uint8_t* dst, void* stream);

/* ---- (2) neighbour-sweep operators.  Inputs in sorted order, outputs in original order.
 *      Each replaces one kernel factory of code/sphops/ops/operators_batch.py. */
/* volume_forward (:16-77): v_i = 1/(Cw sum_j (h^2-d^2)^3_+).  Reads xyz of xv4; writes v (original
 * order) and xv4_out[p] = {x,y,z,v} (sorted order; a buffer distinct from xv4). */
int32_t sphnca_volume_fwd(const sphnca_grid* g, const float* xv4, float* v, float* xv4_out, void* stream);
/* Volumes of the ghost copies of a periodic plane (see sphnca_grid.image_volume): xv4 = {x,y,z,v} with the particles' own
 * (periodic) volumes, image_volume [N][9] out.  g->image_volume is ignored here; band_lo / band_hi / period are read. */
int32_t sphnca_volume_images(const sphnca_grid* g, const float* xv4, float* image_volume, void* stream);
/* volume_backward (:79-151): dLdx [N,D] original order; dLdv_s sorted order */
int32_t sphnca_volume_bwd(const sphnca_grid* g, const float* xv4, const float* dLdv_s, float* dLdx, void* stream);
/* count_forward (:536-597) */
int32_t sphnca_count_fwd(const sphnca_grid* g, const float* xv4, float* c, void* stream);
/* gradient_forward (:158-235): A_s [N,F] sorted -> GA [N,F,D] original order */
int32_t sphnca_gradient_fwd(const sphnca_grid* g, const float* xv4, const float* A_s, int32_t F, float* GA, void* stream);
/* gradient_backward, dL/dA branch (:237-343): G_s = dL/dGA [N,F,D] sorted -> dLdA [N,F] original order */
int32_t sphnca_gradient_bwd(const sphnca_grid* g, const float* xv4, const float* G_s, int32_t F, float* dLdA, void* stream);
/* blur_forward (:600-671): A_s [N,F] -> SA [N,F] */
int32_t sphnca_blur_fwd(const sphnca_grid* g, const float* xv4, const float* A_s, int32_t F, float* SA, void* stream);
/* adjoint of blur w.r.t. A (true adjoint; the reference's :673-772 is not) */
int32_t sphnca_blur_bwd(const sphnca_grid* g, const float* xv4, const float* G_s, int32_t F, float* dLdA, void* stream);
/* divergence_forward (:346-423): A_s [N,F,D] -> DA [N,F] */
int32_t sphnca_divergence_fwd(const sphnca_grid* g, const float* xv4, const float* A_s, int32_t F, float* DA, void* stream);
/* adjoint of divergence w.r.t. A (true adjoint; reference :425-533 indexes the wrong axis) */
int32_t sphnca_divergence_bwd(const sphnca_grid* g, const float* xv4, const float* G_s, int32_t F, float* dLdA, void* stream);

/* ---- (3) fused NCA step (code/nca.py:87-117 with the default feature process), 16 channels,
 *      hidden 256.  Everything in sorted order.  See DESIGN.md 5 for the kernel split. */
typedef struct sphnca_step_params {
    const float* W1;  /* [256, 48] row-major == model.0.weight */
    const float* b1;  /* [256] */
    const float* W2;  /* [n_out, 256] == model.2.weight; n_out = 33 (gated) or 16 (orig) */
    const float* b2;  /* [n_out] */
    int32_t update_rule;   /* 0 = gated, 1 = orig (nca.py:102-108) */
    int32_t use_alpha;     /* nca.py:17-21 */
    float perception_scale;/* h*normalize_perception if >0 else 1 (nca.py:65-69) */
    float orig_rate;       /* self.fire_rate / fire_rate, 'orig' rule only (nca.py:108) */
    /* Layout of the dL/dY rows (dY_s) that sphnca_step_bwd_list WRITES and sphnca_step_adjoint_pairs[_fired] READ -- the two ends
     * of the fire-list backward.  0: the perception vector's own layout, [dYx (16) | dYy (16)].  1: per 4-channel quarter q the
     * pair [dYx 4q..4q+3 | dYy 4q..4q+3]: the 4 lanes that gather a neighbour's row in the adjoint sweep then take their x and y
     * quarters with ONE 32-byte load each (one L1 line lookup per row instead of two; same arithmetic, bit-identical results).
     * Every other entry (sphnca_step_bwd, sphnca_step_adjoint, sphnca_step_bwd_unfired: zeros) uses layout 0 and ignores it. */
    int32_t dy_interleaved;
} sphnca_step_params;

/* One forward step = sphnca_step_fwd then sphnca_step_mask (one kernel each).
 *
 * sphnca_step_fwd: perception sweep (SPH gradient x,y of 16 channels + previous life mask) fused
 * with the 48->256->n_out MLP, the gated/orig update and the fire mask.
 *   in : xv4, A_s [N,16], fire_s uint8 [N] (the `torch.rand(N) <= fire_rate` draw)
 *   out: A_new_s [N,16] (not yet life-masked), alive_s uint8 [N] (activity(A_new) > 0.1),
 *        saved for backward: Y_s [N,32] (scaled dA/dx | dA/dy), O_s [N,36] (MLP output before the
 *        gate non-linearities, columns >= n_out zero), flags_s uint8 [N] (bit0 prev_mask, bit2 fire),
 *        relu_s uint32 [N,8] (bit n%32 of word n/32 = hidden unit n is active, i.e. pre-activation > 0). */
int32_t sphnca_step_fwd(const sphnca_grid* g, const float* xv4, const float* A_s,
                        const uint8_t* fire_s, const sphnca_step_params* p,
                        float* A_new_s, uint8_t* alive_s, float* Y_s, float* O_s, uint8_t* flags_s,
                        uint32_t* relu_s, void* stream);
/* The two halves of the default (split) forward, callable on their own:
 *   sphnca_step_perceive  neighbour sweep: Y_s [N,32] (scaled dA/dx | dA/dy), flags_s bit0 = previous life mask;
 *   sphnca_step_mlp       tcgen05 MLP + gated/orig update + fire mask from (A_s, Y_s); reads flags_s bit0,
 *                         writes A_new_s, alive_s, O_s, flags_s (bit0 | fire << 2), relu_s. */
int32_t sphnca_step_perceive(const sphnca_grid* g, const float* xv4, const float* A_s, const sphnca_step_params* p,
                             float* Y_s, uint8_t* flags_s, void* stream);
int32_t sphnca_step_mlp(int64_t N, const float* A_s, const float* Y_s, const uint8_t* fire_s, const sphnca_step_params* p,
                        float* A_new_s, uint8_t* alive_s, float* O_s, uint8_t* flags_s, uint32_t* relu_s, void* stream);
/* Implementation switch of the step kernels (A/B tests; all produce the same tensors to fp32 rounding):
 *   1 (default) tcgen05 MLP with the 3xTF32 split; forward split into a full-L1 perception sweep
 *               kernel + the tensor-core MLP kernel (csrc/step.cu:step_perceive_kernel, step_mlp_tc.cu);
 *   2           same, forward fused into one persistent kernel (csrc/step_tc.cu);
 *   0           fp32 CUDA-core MLP (csrc/step.cu). */
int32_t sphnca_set_mlp_path(int32_t path);
int32_t sphnca_get_mlp_path(void);
/* Implementation switch of the two hot sweeps (sphnca_step_perceive, sphnca_step_adjoint): 0 (default) =
 * one-thread-per-particle gathers out of L1/L2 with packed-pair FMAs, 1 = shared-memory tile engine with
 * in-radius compaction (csrc/sweep_tile.cuh; measured slower on B200, kept for A/B), 2 = sphnca_step_perceive
 * runs the TMA-staged kernel below. */
int32_t sphnca_set_sweep_path(int32_t path);
/* sphnca_step_perceive with the candidates of every 128-particle tile staged in shared memory by TMA bulk copies
 * (csrc/sweep_stage.cuh).  Same arguments, bit-identical results.  Faster when (nearly) all tiles fit the 44 KB staging
 * buffer -- planar / padded-2-D clouds, ~0.27 vs 0.30 ms on BASELINE configs[1] -- and slower when they do not (3-D
 * clouds: every tile plans, gives up and gathers at a lower occupancy).  sphnca_stage_fit counts both:
 * counts[0] = tiles, counts[1] = tiles that do not fit (device int32[2], asynchronous on `stream`). */
int32_t sphnca_step_perceive_staged(const sphnca_grid* g, const float* xv4, const float* A_s, const sphnca_step_params* p,
                                    float* Y_s, uint8_t* flags_s, void* stream);
int32_t sphnca_stage_fit(const sphnca_grid* g, int32_t* counts, void* stream);

/* Pair lists (csrc/pairs.cu).  Positions and volumes are fixed over the K steps of a rollout (nca.py:141-145), so the
 * 3^D-cell walk, the in-radius test and the kernel weights of the two hot sweeps are computed once per grid and
 * stored per in-radius pair, [slot][8] per group of 8 consecutive sorted particles (a group's slot count = the largest
 * neighbour count of its particles rounded up to a multiple of 4; shorter lists are padded with zero-weight self pairs):
 *   sphnca_pairs_count   group_slots int32 [ceil(N/8)]
 *   caller               group_base int64 [ceil(N/8)] = 8 * exclusive prefix sum of group_slots; E = 8 * sum
 *   sphnca_pairs_fill    J int32 [E], WP / WB / WA float [E,2], SXY float [N,2], FLG uint8 [N] (the life mask of an
 *                        all-alive field: what the perception sweep reports when use_alpha == 0); WB (the life-mask blur
 *                        weights, 8 of the 28 bytes per pair) may be NULL when no model with an alpha channel will use the
 *                        lists -- sphnca_step_perceive_pairs then requires use_alpha == 0; WPZ float [E] (optional, NULL = not
 *                        built): the z component of the perception weight, read by sphnca_step_perceive_pairs_frame only
 *   sphnca_step_perceive_pairs / sphnca_step_adjoint_pairs: same outputs, bit for bit, as sphnca_step_perceive /
 *   sphnca_step_adjoint (same visiting order, same floats), from the lists. */
int32_t sphnca_pairs_count(const sphnca_grid* g, const float* xv4, int32_t* group_slots, void* stream);
int32_t sphnca_pairs_fill(const sphnca_grid* g, const float* xv4, const int32_t* group_slots, const int64_t* group_base,
                          int32_t* J, float* WP, float* WB, float* WA, float* SXY, uint8_t* FLG, float* WPZ, void* stream);
/* Life mask of a model WITH an alpha channel from the pair lists: the F = 1 blur of nca.py:71-79 alone (12 list bytes + one alive
 * byte per neighbour), so that the 16-channel perception can run over the fired rows only.  mode 0: flags_s = bit 0 (previous
 * mask; alive_in = alive bytes of the current state).  mode 1: flags_s |= bit 1 (new mask; alive_in = alive bytes of the unmasked
 * new state), A_new_s rows with !(prev & new) are multiplied by 0, alive_out (optional) = alive bytes of the masked new state.
 * Bit-identical to the masks of sphnca_step_perceive_pairs / sphnca_step_mask. */
int32_t sphnca_step_mask_pairs(const sphnca_grid* g, const int32_t* group_slots, const int64_t* group_base, const int32_t* J,
                               const float* WB, const uint8_t* alive_in, int32_t mode, uint8_t* flags_s, float* A_new_s,
                               uint8_t* alive_out, void* stream);
/* Perception in a per-particle tangent frame (surface clouds: reference nca.py:325-336 `project_tangent_space` applied to the
 * gradient of every channel, then `[..., 0]` and `[..., 1]` kept by the default feature process): Y_s = [A-gradient along T_i |
 * along B_i], frame_s [N][8] = {T_i, 0, B_i = N_i x T_i, 0} in cell-sorted order.  The projection is applied to the pair weights
 * (w . T_i, w . B_i), so the sweep costs what the planar one costs and feeds the same MLP kernel. */
int32_t sphnca_step_perceive_pairs_frame(const sphnca_grid* g, const int32_t* group_slots, const int64_t* group_base, const int32_t* J,
                                         const float* WP, const float* WPZ, const float* WB, const uint8_t* FLG, const float* frame_s,
                                         const float* A_s, const sphnca_step_params* p, float* Y_s, uint8_t* flags_s, void* stream);
int32_t sphnca_step_perceive_pairs(const sphnca_grid* g, const int32_t* group_slots, const int64_t* group_base, const int32_t* J,
                                   const float* WP, const float* WB, const uint8_t* FLG, const float* A_s,
                                   const sphnca_step_params* p, float* Y_s, uint8_t* flags_s, void* stream);
int32_t sphnca_step_adjoint_pairs(const sphnca_grid* g, const int32_t* group_slots, const int64_t* group_base, const int32_t* J,
                                  const float* WA, const float* SXY, const float* xv4, const float* dY_s, const float* dAdir_s,
                                  const sphnca_step_params* p, float* dA_s, void* stream);
/* sphnca_step_mask: new life mask = blur(alive_s) > 0.1 (nca.py:112); A_new_s *= prev & new in
 * place; flags_s bit1 = new_mask. */
int32_t sphnca_step_mask(const sphnca_grid* g, const float* xv4, const uint8_t* alive_s,
                         float* A_new_s, uint8_t* flags_s, void* stream);
/* The same update when use_alpha == 0 (reference nca.py:17-21: the activity field is all-ones, so the new life
 * mask is the blur of the same constant field as the previous one, i.e. bit0 of flags_s as written by the perception
 * sweep): flags_s bit1 = bit0, rows with bit0 == 0 are multiplied by 0.  Elementwise, no neighbour sweep. */
int32_t sphnca_step_mask_static(int64_t N, float* A_new_s, uint8_t* flags_s, void* stream);

/* One backward step = sphnca_step_bwd then sphnca_step_adjoint (one kernel each), gather-based,
 * no float atomics.
 * sphnca_step_bwd: everything after the perception sweep.  dA_next_s [N,16] ->
 *   dO_s [N,36] = dL/d(MLP output) (for the weight-gradient stage), dY_s [N,32] = dL/dY,
 *   dAdir_s [N,16] = the part of dL/dA that does not pass through the sweep. */
int32_t sphnca_step_bwd(int64_t N, const float* A_s, const float* Y_s, const float* O_s,
                        const uint8_t* flags_s, const uint32_t* relu_s, const float* dA_next_s,
                        const sphnca_step_params* p, float* dO_s, float* dY_s, float* dAdir_s,
                        float* Ht, float* dHt, void* stream);
/*   relu_s: the ReLU bits saved by sphnca_step_fwd (used by the tensor-core path, which needs no recompute). */
/*   Ht, dHt (both or neither, may be NULL): hidden activations relu(pre) and dL/d(pre), transposed and
 *   tile-blocked for sphnca_mlp_wgrad: float [ceil(N/32)][256][32], element (p, n) at
 *   ((p/32)*256 + n)*32 + p%32. */

/* Weight gradients of the MLP, ACCUMULATED (+=) into dW1 [256,48], db1 [256], dW2 [n_out,256],
 * db2 [n_out] (replaces torch autograd's GEMMs over `model.0` / `model.2`, code/nca.py:96,151).
 * Deterministic two-stage reduction.  With Ht == dHt == NULL (default) the hidden layer is recomputed
 * on the tensor cores from (A_s, Y_s, dO_s) and the weights in `p`; with Ht/dHt given (the CUDA-core
 * sphnca_step_bwd wrote them) they are streamed instead. */
int32_t sphnca_mlp_wgrad_workspace_bytes(int64_t N, size_t* bytes);
int32_t sphnca_mlp_wgrad(int64_t N, const float* A_s, const float* Y_s, const float* dO_s, const float* Ht,
                         const float* dHt, const sphnca_step_params* p, float* dW1, float* db1, float* dW2,
                         float* db2, void* workspace, size_t workspace_bytes, void* stream);
/* sphnca_step_adjoint: dA_s = dAdir_s + (adjoint of the perception sweep applied to dY_s). */
int32_t sphnca_step_adjoint(const sphnca_grid* g, const float* xv4, const float* dY_s, const float* dAdir_s,
                            const sphnca_step_params* p, float* dA_s, void* stream);

/* ---- fire lists: the MLP half of a step only for the particles whose fire-mask draw came up ------------------------
 * Reference code/nca.py:100-107 evaluates the MLP for every particle and keeps the old state where the stochastic
 * update mask is off (`torch.where(update_mask, nA, A)`): for those particles the MLP output is discarded and no
 * gradient reaches the MLP, so skipping them changes no result.  With fire rate 0.5 it halves the tensor-core work.
 *   sphnca_fire_list        rows[0..*count) = sorted positions of the fired particles in increasing order (deterministic
 *                           ordered compaction); not-fired rows: A_new_s = A_s, alive byte.  count is a DEVICE int32 that
 *                           the _list kernels read themselves: no host synchronisation.  A_new_s = alive_s = NULL: the
 *                           list only, no row is carried over (a second list over the same step, multigpu.py).
 *   sphnca_step_mlp_list    sphnca_step_mlp over the listed rows.  O_c [N,36], relu_c [N,8] and the optional compact
 *                           copies A_c [N,16], Y_c [N,32] (contiguous input of sphnca_mlp_wgrad_list) are indexed by
 *                           list position; A_new_s, alive_s, flags_s by particle.  inputs_by_list != 0: A_s / Y_s are
 *                           such compact arrays already (written by sphnca_step_perceive_pairs_fired), A_c = Y_c = NULL.
 *   sphnca_step_perceive_pairs_fired  the perception sweep for the listed particles only (the perception vector feeds nothing
 *                           but the MLP; with an alpha channel the life mask comes from sphnca_step_mask_pairs); writes Y_c and the compact state rows A_c by
 *                           list position, bit-identical to the rows of the full sweep.  sphnca_fire_list's FLG / flags_s
 *                           pair (optional, both or neither) copies the static life-mask bit of every row meanwhile.
 *   sphnca_step_bwd_list    sphnca_step_bwd over the listed rows (dO_c by list position, dY_s / dAdir_s by particle);
 *   sphnca_step_bwd_unfired the not-fired rows of the same step: dY_s = 0, dAdir_s = living ? dA_next_s : 0.
 *   sphnca_mlp_wgrad_list   sphnca_mlp_wgrad over the first *count rows of the compact arrays (workspace sized for N).
 * The fire-list workspace (sphnca_fire_list_workspace_bytes, 8-byte aligned) holds the look-back state of the one-launch ordered
 * compaction: it must be ZERO before its first use; every call leaves it zeroed again (calls on one workspace must not overlap). */
int32_t sphnca_fire_list_workspace_bytes(int64_t N, size_t* bytes);
int32_t sphnca_fire_list(int64_t N, const uint8_t* fire_s, const float* A_s, int32_t use_alpha, int32_t* rows,
                         int32_t* count, float* A_new_s, uint8_t* alive_s, const uint8_t* FLG, uint8_t* flags_s,
                         int32_t fold_static_mask, void* workspace, size_t workspace_bytes, void* stream);
/* sphnca_fire_list with the draw fused in: fire byte of sorted row p = (u[sorted_id[p]] <= rate), u = the reference's
 * `torch.rand(N)` in particle order (nca.py:100); fire_s [N] receives the bytes (kept for activation checkpointing).
 * fold_static_mask != 0 (both entries; no alpha channel, FLG given): the life mask of the step is applied right here and by
 * sphnca_step_mlp_list(inputs_by_list | 2) -- flag bits 0|1 set for living rows, dead rows multiplied by 0 -- so that
 * no sphnca_step_mask_static pass is needed. */
int32_t sphnca_fire_list_draw(const sphnca_grid* g, const float* u, float rate, uint8_t* fire_s, const float* A_s,
                              int32_t use_alpha, int32_t* rows, int32_t* count, float* A_new_s, uint8_t* alive_s,
                              const uint8_t* FLG, uint8_t* flags_s, int32_t fold_static_mask, void* workspace,
                              size_t workspace_bytes, void* stream);
int32_t sphnca_step_perceive_pairs_fired(const sphnca_grid* g, const int32_t* rows, const int32_t* count, const int32_t* group_slots,
                                         const int64_t* group_base, const int32_t* J, const float* WP, const float* A_s,
                                         const sphnca_step_params* p, float* Y_c, float* A_c, void* stream);
int32_t sphnca_step_mlp_list(int64_t N, const int32_t* rows, const int32_t* count, const float* A_s, const float* Y_s,
                             const sphnca_step_params* p, float* A_new_s, uint8_t* alive_s, float* O_c, uint8_t* flags_s,
                             uint32_t* relu_c, float* A_c, float* Y_c, int32_t inputs_by_list, void* stream);
int32_t sphnca_step_bwd_list(int64_t N, const int32_t* rows, const int32_t* count, const float* A_c, const float* O_c,
                             const uint8_t* flags_s, const uint32_t* relu_c, const float* dA_next_s,
                             const sphnca_step_params* p, float* dO_c, float* dY_s, float* dAdir_s, void* stream);
int32_t sphnca_step_bwd_unfired(int64_t N, const uint8_t* flags_s, const float* dA_next_s, float* dY_s, float* dAdir_s,
                                void* stream);
/* The whole forward of a no-alpha fire-list step in ONE call (host overhead: 1 instead of 3 C-ABI calls): sphnca_fire_list
 * (fold_static_mask = 1) -> sphnca_step_perceive_pairs_fired -> sphnca_step_mlp_list(inputs_by_list | fold).  A step may be split
 * into several lists (multigpu.py: boundary particles first, interior particles while the halo exchange runs): carry_sel [N]
 * (NULL = all) selects the rows THIS call carries over when they did not fire, write_flags = 1 on the first list of a step (the
 * static mask bits of all rows), max_fired = an upper bound of this list's length when the caller knows one (0 = N; sizes the
 * grids of the list kernels). */
int32_t sphnca_step_fwd_fired(const sphnca_grid* g, const uint8_t* fire_s, const float* A_s, const int32_t* group_slots,
                              const int64_t* group_base, const int32_t* J, const float* WP, const uint8_t* FLG,
                              const sphnca_step_params* p, const uint8_t* carry_sel, int32_t write_flags, int64_t max_fired,
                              int32_t* rows, int32_t* count, float* Y_c, float* A_c, float* A_new_s, uint8_t* alive_s, float* O_c,
                              uint8_t* flags_s, uint32_t* relu_c, void* workspace, size_t workspace_bytes, void* stream);
/* Counter-based fire bytes of one step for slab-sharded rollouts (multigpu.py; no reference counterpart -- the reference draws
 * torch.rand(N) on its single GPU, nca.py:100): fire_s[p] = gid_s[p] >= 0 && U(gid_s[p], seed) <= rate, U = top 24 bits of
 * splitmix64(gid + 0x9E3779B97F4A7C15 (seed + 1)) / 2^24 -- a function of the GLOBAL particle id, so N ranks draw what 1 rank
 * draws.  gid_s [N] int64 in cell-sorted order, -1 = halo particle (never fires).  Optional: bnd_s [N] (1 = boundary particle)
 * -> fire_b = fire & bnd, fire_i = fire & ~bnd (the two passes of a split step). */
int32_t sphnca_fire_bytes_counter(int64_t N, const int64_t* gid_s, int64_t seed, float rate, const uint8_t* bnd_s, uint8_t* fire_s,
                                  uint8_t* fire_b, uint8_t* fire_i, void* stream);
/* sphnca_step_adjoint_pairs for a fire-list step: rows of dY_s / dAdir_s are read only for the particles whose flag byte has the
 * fire bit (4); for the others dL/dY is zero and the direct part is (living ? dA_next_s : 0), taken from flags_s / dA_next_s --
 * no sphnca_step_bwd_unfired pass is needed and the row gathers of not-fired neighbours are skipped. */
int32_t sphnca_step_adjoint_pairs_fired(const sphnca_grid* g, const int32_t* group_slots, const int64_t* group_base, const int32_t* J,
                                        const float* WA, const float* SXY, const float* xv4, const float* dY_s, const float* dAdir_s,
                                        const uint8_t* flags_s, const float* dA_next_s, const sphnca_step_params* p, float* dA_s,
                                        void* stream);
int32_t sphnca_mlp_wgrad_list(int64_t N, const int32_t* count, const float* A_c, const float* Y_c, const float* dO_c,
                              const sphnca_step_params* p, float* dW1, float* db1, float* dW2, float* db2,
                              void* workspace, size_t workspace_bytes, void* stream);
/* sphnca_mlp_wgrad_list in two halves, for the backward of a whole rollout (replaces per-step Linear.backward accumulation into
 * .grad, code/nca.py:62-79 through torch autograd): every step's call ADDS its rows' contribution to the per-stream partial sums
 * in `workspace` (accumulate = 0 on the first call of a series: plain stores), sphnca_mlp_wgrad_reduce then adds the partials
 * to the parameter gradients ONCE, in a fixed order.  All calls of a series use the same N (it fixes the stream count) and the
 * same workspace (sphnca_mlp_wgrad_workspace_bytes(N)). */
int32_t sphnca_mlp_wgrad_list_partial(int64_t N, const int32_t* count, const float* A_c, const float* Y_c, const float* dO_c,
                                      const sphnca_step_params* p, int32_t accumulate, void* workspace, size_t workspace_bytes,
                                      void* stream);
int32_t sphnca_mlp_wgrad_reduce(int64_t N, const sphnca_step_params* p, float* dW1, float* db1, float* dW2, float* db2,
                                const void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPHNCA_H */
