"""Fused NCA step (B200 extension of the `sphops` package; no counterpart module in the reference).

`nca_step` computes one `SPHNCA.forward` (reference code/nca.py:87-117, default feature process,
16 channels, hidden width 256) with three kernels instead of ~30 launches:
perception sweep (+ previous life mask), MLP + gated update, life-mask sweep (+ masking).
Backward: MLP backward (+ parameter gradients, deterministic) and the gather adjoint of the
perception sweep.  All state is handled in cell-sorted order inside; inputs/outputs are in the
caller's original particle order like every other sphops operator.

`nca_rollout` runs K such steps as ONE autograd node with the state resident in cell-sorted order
(`_NCARollout`): per step a fire list (the particles whose fire draw came up), the perception sweep
and the MLP over those particles only, activation checkpointing on request -- the default path of
`nca.sample_plane` / `SPHNCA.rollout` and of bench.py.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, ptr, require_cuda_f32, stream_ptr

_RULES = {"gated": 0, "orig": 1}


def _step_params(W1, b1, W2, b2, update_rule, use_alpha, perception_scale, orig_rate):
    p = _lib.CStepParams()
    p.W1, p.b1, p.W2, p.b2 = W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr()
    p.update_rule = _RULES[update_rule]
    p.use_alpha = int(bool(use_alpha))
    p.perception_scale = float(perception_scale)
    p.orig_rate = float(orig_rate)
    return p


# Optional tape probe used by the parity tests: a list that receives, per taped (checkpoint=False) fire-list step of
# `_NCARollout.forward`, the tuple (rows, count, relu_c, Y_c, O_c) -- the fired rows' sorted positions, their device-side
# count, their 256 saved ReLU decisions (8 words per list position; bit b of word w = hidden unit 32 w + b), their perception
# rows and MLP outputs -- so that the oracle can be evaluated with the GPU's own ReLU decisions imposed
# (oracle.nca_step(relu_bits=...)) and the stages of a step compared one by one.  None = off (default).
TAPE_PROBE = None

# Optional per-kernel timing hook used by bench.py: a dict name -> list of (start, end) CUDA events
# recorded on the launching (current) stream around each C-ABI call.  None = off (default).
PROFILE = None


def _timed(name, fn, *args):
    if PROFILE is None:
        check(fn(*args))
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    check(fn(*args))
    b.record()
    PROFILE.setdefault(name, []).append((a, b))


def _perceive_entry(grid):
    """The TMA-staged perception sweep when the grid's tiles fit its staging buffer, else the gather sweep."""
    return lib.sphnca_step_perceive_staged if grid.staged_sweep_ok() else lib.sphnca_step_perceive


def _perceive(grid, x, v, g, xv4, A_s, p, Y_s, flags_s, st):
    """Perception sweep: from the grid's pair lists when they exist, else by sweeping the cells."""
    pl = grid.pair_lists(x, v, blur=bool(p.use_alpha))
    _perceive_pl(grid, pl, g, xv4, A_s, p, Y_s, flags_s, st)
    return pl


def _perceive_pl(grid, pl, g, xv4, A_s, p, Y_s, flags_s, st):
    """Perception sweep with the pair lists (or their absence) already looked up."""
    if pl is not None:
        _timed("step_perceive", lib.sphnca_step_perceive_pairs, g, ptr(pl["slots"]), ptr(pl["base"]), ptr(pl["J"]), ptr(pl["WP"]),
               ptr(pl["WB"]), ptr(pl["FLG"]), ptr(A_s), C.byref(p), ptr(Y_s), ptr(flags_s), st)
    else:
        _timed("step_perceive", _perceive_entry(grid), g, ptr(xv4), ptr(A_s), C.byref(p), ptr(Y_s), ptr(flags_s), st)


def _adjoint(grid, pl, xv4, dY_s, dAdir_s, p, dA_s, st):
    if pl is not None:
        _timed("step_adjoint", lib.sphnca_step_adjoint_pairs, grid.c_grid(), ptr(pl["slots"]), ptr(pl["base"]), ptr(pl["J"]),
               ptr(pl["WA"]), ptr(pl["SXY"]), ptr(xv4), ptr(dY_s), ptr(dAdir_s), C.byref(p), ptr(dA_s), st)
    else:
        _timed("step_adjoint", lib.sphnca_step_adjoint, grid.c_grid(), ptr(xv4), ptr(dY_s), ptr(dAdir_s), C.byref(p), ptr(dA_s), st)


def _mask(use_alpha, g, N, xv4, alive_s, An_s, flags_s, st):
    """Life-mask update of a step.  Without an alpha channel the new mask equals the previous one (both are the blur
    of an all-ones field, reference nca.py:17-21,94-98), so no second neighbour sweep is needed."""
    if use_alpha:
        _timed("step_mask", lib.sphnca_step_mask, g, ptr(xv4), ptr(alive_s), ptr(An_s), ptr(flags_s), st)
    else:
        _timed("step_mask_static", lib.sphnca_step_mask_static, N, ptr(An_s), ptr(flags_s), st)


class _NCAStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, v, A, update_mask, W1, b1, W2, b2, h, grid, update_rule, use_alpha, perception_scale, orig_rate, frame=None):
        x = require_cuda_f32(x, "x"); v = require_cuda_f32(v, "v"); A_ = require_cuda_f32(A, "A")
        W1c, b1c, W2c, b2c = [require_cuda_f32(t, n) for t, n in ((W1, "W1"), (b1, "b1"), (W2, "W2"), (b2, "b2"))]
        N = x.shape[0]
        n_out = 33 if update_rule == "gated" else 16
        if A_.shape != (N, 16) or W1c.shape != (256, 48) or W2c.shape != (n_out, 256):
            raise ValueError("fused step needs A [N,16], W1 [256,48], W2 [33|16,256]")
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        xv4 = grid.sorted_xv(x, v)
        A_s = grid.pack_rows(A_, 16)
        fire_s = grid.pack_bytes(update_mask.to(torch.uint8).contiguous())
        An_s = torch.empty((N, 16), **f32)
        Y_s = torch.empty((N, 32), **f32)
        O_s = torch.empty((N, 36), **f32)
        flags_s = torch.empty((N,), dtype=torch.uint8, device=dev)
        alive_s = torch.empty((N,), dtype=torch.uint8, device=dev)
        relu_s = torch.empty((N, 8), dtype=torch.int32, device=dev)
        p = _step_params(W1c, b1c, W2c, b2c, update_rule, use_alpha, perception_scale, orig_rate)
        if N:
            g, st = grid.c_grid(), stream_ptr()
            if frame is not None:
                # perception projected onto the particles' tangent frames inside the sweep (inference only; surface clouds)
                pl = grid.pair_lists(x, v, blur=bool(use_alpha), z=True)
                frame_s = grid.pack_rows(frame, 8)
                _timed("step_perceive", lib.sphnca_step_perceive_pairs_frame, g, ptr(pl["slots"]), ptr(pl["base"]), ptr(pl["J"]), ptr(pl["WP"]),
                       ptr(pl["WPZ"]), ptr(pl["WB"]), ptr(pl["FLG"]), ptr(frame_s), ptr(A_s), C.byref(p), ptr(Y_s), ptr(flags_s), st)
                _timed("step_mlp", lib.sphnca_step_mlp, N, ptr(A_s), ptr(Y_s), ptr(fire_s), C.byref(p), ptr(An_s),
                       ptr(alive_s), ptr(O_s), ptr(flags_s), ptr(relu_s), st)
            elif SPLIT_FORWARD:
                _perceive(grid, x, v, g, xv4, A_s, p, Y_s, flags_s, st)
                _timed("step_mlp", lib.sphnca_step_mlp, N, ptr(A_s), ptr(Y_s), ptr(fire_s), C.byref(p), ptr(An_s),
                       ptr(alive_s), ptr(O_s), ptr(flags_s), ptr(relu_s), st)
            else:
                _timed("step_fwd", lib.sphnca_step_fwd, g, ptr(xv4), ptr(A_s), ptr(fire_s), C.byref(p), ptr(An_s),
                       ptr(alive_s), ptr(Y_s), ptr(O_s), ptr(flags_s), ptr(relu_s), st)
            _mask(use_alpha, g, N, xv4, alive_s, An_s, flags_s, st)
        out = grid.unpack_rows(An_s, 16)
        if frame is not None:                        # no tape: the frame sweep has no adjoint (callers fall back when gradients are needed)
            ctx.mark_non_differentiable(out)
            return out
        ctx.grid = grid
        ctx.pairs = grid.pair_lists(x, v, blur=bool(use_alpha)) if N else None     # cached by the grid; the adjoint sweep reads the same lists
        ctx.cfg = (update_rule, use_alpha, perception_scale, orig_rate)
        ctx.save_for_backward(xv4, A_s, Y_s, O_s, flags_s, relu_s, W1c, b1c, W2c, b2c)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dA_next):
        xv4, A_s, Y_s, O_s, flags_s, relu_s, W1, b1, W2, b2 = ctx.saved_tensors
        grid = ctx.grid
        N = A_s.shape[0]
        dev = A_s.device
        f32 = dict(dtype=torch.float32, device=dev)
        n_out = W2.shape[0]
        dAn_s = grid.pack_rows(require_cuda_f32(dA_next, "dA_next"), 16)
        dA_s = torch.empty((N, 16), **f32)
        dO_s = torch.empty((N, 36), **f32)
        dY_s = torch.empty((N, 32), **f32)
        dAdir_s = torch.empty((N, 16), **f32)
        p = _step_params(W1, b1, W2, b2, *ctx.cfg)
        g = ctx.needs_input_grad
        want_w = any(g[4:8])
        dW1 = db1 = dW2 = db2 = None
        if want_w:
            dW1, db1, dW2, db2 = torch.zeros_like(W1), torch.zeros_like(b1), torch.zeros_like(W2), torch.zeros_like(b2)
        if N:
            st = stream_ptr()
            # particles are processed in chunks so that the transposed hidden activations
            # (2 x [chunk/32][256][32] fp32, 2 KB per particle) stay bounded for very large clouds
            simt = SIMT_BACKWARD
            chunk = min(N, WGRAD_CHUNK) if simt else N
            Ht = dHt = ws = None
            if want_w:
                if simt:
                    tiles = (chunk + 31) // 32      # [tiles][256][32], see include/sphnca.h
                    Ht = torch.empty((tiles, 256, 32), **f32)
                    dHt = torch.empty((tiles, 256, 32), **f32)
                nb = C.c_size_t(0)
                check(lib.sphnca_mlp_wgrad_workspace_bytes(chunk, C.byref(nb)))
                ws = torch.empty((nb.value,), dtype=torch.uint8, device=dev)
            for lo in range(0, N, chunk):
                n = min(chunk, N - lo)
                Hc, dHc = Ht, dHt
                _timed("step_bwd", lib.sphnca_step_bwd, n, ptr(A_s[lo:]), ptr(Y_s[lo:]), ptr(O_s[lo:]), ptr(flags_s[lo:]),
                       ptr(relu_s[lo:]), ptr(dAn_s[lo:]), C.byref(p), ptr(dO_s[lo:]), ptr(dY_s[lo:]), ptr(dAdir_s[lo:]),
                       ptr(Hc), ptr(dHc), st)
                if want_w:
                    _timed("mlp_wgrad", lib.sphnca_mlp_wgrad, n, ptr(A_s[lo:]), ptr(Y_s[lo:]), ptr(dO_s[lo:]), ptr(Hc), ptr(dHc),
                           C.byref(p), ptr(dW1), ptr(db1), ptr(dW2), ptr(db2), ptr(ws), nb.value, st)
            _adjoint(grid, ctx.pairs, xv4, dY_s, dAdir_s, p, dA_s, st)
        dA = grid.unpack_rows(dA_s, 16) if g[2] else None
        return (None, None, dA, None, dW1 if g[4] else None, db1 if g[5] else None,
                dW2 if g[6] else None, db2 if g[7] else None, None, None, None, None, None, None, None)


# The MLP half of a rollout step runs over the fired rows only (csrc/firelist.cu) when the tensor-core kernels are the
# active path; False = every row goes through the MLP kernels like in the single-step node.
FIRE_LISTS = True


def _fire_lists_on():
    return (FIRE_LISTS and SPLIT_FORWARD and not SIMT_BACKWARD and "sphnca_fire_list" not in _lib.MISSING_SYMBOLS
            and lib.sphnca_get_mlp_path() >= 1)


def _fire_list_workspace(N, dev):
    nb = C.c_size_t(0)
    check(lib.sphnca_fire_list_workspace_bytes(N, C.byref(nb)))
    # zero before the first use (look-back state words of the one-launch compaction); every call leaves it zeroed again
    return torch.zeros((nb.value,), dtype=torch.uint8, device=dev)


class _FireBuffers:
    """Per-step tensors of the fire-list path: the list (rows, device-side count), the saved MLP output / ReLU bits and
    the compact copies of the MLP input rows, all indexed by list position, plus the per-particle flag bytes.  Sized for
    N rows: the count stays on the device, so the host never learns how many particles fired."""
    PER_STEP = 7

    def __init__(self, N, dev, compact_inputs=True):
        f32 = dict(dtype=torch.float32, device=dev)
        self.rows = torch.empty((N,), dtype=torch.int32, device=dev)
        self.count = torch.empty((1,), dtype=torch.int32, device=dev)
        self.O_c = torch.empty((N, 36), **f32)
        self.relu_c = torch.empty((N, 8), dtype=torch.int32, device=dev)
        self.flags = torch.empty((N,), dtype=torch.uint8, device=dev)
        self.A_c = torch.empty((N, 16), **f32) if compact_inputs else None
        self.Y_c = torch.empty((N, 32), **f32) if compact_inputs else None

    def tensors(self):
        return [self.rows, self.count, self.A_c, self.Y_c, self.O_c, self.flags, self.relu_c]

    @classmethod
    def from_tensors(cls, t):
        self = cls.__new__(cls)
        self.rows, self.count, self.A_c, self.Y_c, self.O_c, self.flags, self.relu_c = t
        return self


# With an alpha channel: the two life masks of a step from their own light list sweeps (sphnca_step_mask_pairs) and the 16-channel
# perception over the fired rows only, instead of the full perception sweep (+ blur) over all rows and the cell-sweep mask.
ALPHA_MASK_LISTS = __import__("os").environ.get("SPHNCA_ALPHA_MASK_LISTS", "1") != "0"


def _step_forward_lists(grid, x, v, g, xv4, S, fire_s, p, use_alpha, Y_s, Sn, alive_s, fb, ws_fl, st, pl=False, draw=None, alive_io=None):
    """One forward step on the fire-list path: list + carried-over rows, perception sweep, MLP over the list, life mask.
    Without an alpha channel and with pair lists the perception sweep itself runs over the fired particles only and writes the
    compact [A_c | Y_c] rows the MLP kernel reads (needs fb.A_c / fb.Y_c); otherwise the full sweep fills Y_s.
    draw = (u, rate): the fire bytes are made right here from the reference's `torch.rand(N)` values u (particle order) as
    u <= rate and written to fire_s; on the fired-sweep path the static life mask is then applied by the list and MLP kernels
    themselves (no separate mask pass)."""
    N = S.shape[0]
    if pl is False:
        pl = grid.pair_lists(x, v, blur=bool(use_alpha))
    if use_alpha and ALPHA_MASK_LISTS and pl is not None and pl["WB"] is not None and fb.A_c is not None \
            and "sphnca_step_mask_pairs" not in _lib.MISSING_SYMBOLS:
        # alive_io = (alive bytes of S, buffer for the alive bytes of the masked new state): chained by the forward loop; a lone
        # step (recompute in the backward) derives the former from the state
        alive_cur, alive_next = alive_io if alive_io is not None else ((S[:, 3] > 0.1).view(torch.uint8).contiguous(), None)
        lists = (ptr(pl["slots"]), ptr(pl["base"]), ptr(pl["J"]))
        _timed("step_mask", lib.sphnca_step_mask_pairs, g, *lists, ptr(pl["WB"]), ptr(alive_cur), 0, ptr(fb.flags), None, None, st)
        if draw is not None:
            _timed("fire_list", lib.sphnca_fire_list_draw, g, ptr(draw[0]), float(draw[1]), ptr(fire_s), ptr(S), 1, ptr(fb.rows), ptr(fb.count),
                   ptr(Sn), ptr(alive_s), None, None, 0, ptr(ws_fl), ws_fl.numel(), st)
        else:
            _timed("fire_list", lib.sphnca_fire_list, N, ptr(fire_s), ptr(S), 1, ptr(fb.rows), ptr(fb.count), ptr(Sn), ptr(alive_s), None, None, 0,
                   ptr(ws_fl), ws_fl.numel(), st)
        _timed("step_perceive", lib.sphnca_step_perceive_pairs_fired, g, ptr(fb.rows), ptr(fb.count), *lists, ptr(pl["WP"]), ptr(S), C.byref(p),
               ptr(fb.Y_c), ptr(fb.A_c), st)
        _timed("step_mlp", lib.sphnca_step_mlp_list, N, ptr(fb.rows), ptr(fb.count), ptr(fb.A_c), ptr(fb.Y_c), C.byref(p), ptr(Sn), ptr(alive_s),
               ptr(fb.O_c), ptr(fb.flags), ptr(fb.relu_c), None, None, 1, st)
        _timed("step_mask", lib.sphnca_step_mask_pairs, g, *lists, ptr(pl["WB"]), ptr(alive_s), 1, ptr(fb.flags), ptr(Sn), ptr(alive_next), st)
        return
    fired_sweep = pl is not None and not use_alpha and fb.A_c is not None
    fold = fired_sweep
    flg, flags = (ptr(pl["FLG"]), ptr(fb.flags)) if fired_sweep else (None, None)
    if draw is not None:
        _timed("fire_list", lib.sphnca_fire_list_draw, g, ptr(draw[0]), float(draw[1]), ptr(fire_s), ptr(S), int(bool(use_alpha)),
               ptr(fb.rows), ptr(fb.count), ptr(Sn), ptr(alive_s), flg, flags, int(fold), ptr(ws_fl), ws_fl.numel(), st)
    else:
        _timed("fire_list", lib.sphnca_fire_list, N, ptr(fire_s), ptr(S), int(bool(use_alpha)), ptr(fb.rows), ptr(fb.count), ptr(Sn),
               ptr(alive_s), flg, flags, int(fold), ptr(ws_fl), ws_fl.numel(), st)
    if fired_sweep:
        _timed("step_perceive", lib.sphnca_step_perceive_pairs_fired, g, ptr(fb.rows), ptr(fb.count), ptr(pl["slots"]), ptr(pl["base"]),
               ptr(pl["J"]), ptr(pl["WP"]), ptr(S), C.byref(p), ptr(fb.Y_c), ptr(fb.A_c), st)
        _timed("step_mlp", lib.sphnca_step_mlp_list, N, ptr(fb.rows), ptr(fb.count), ptr(fb.A_c), ptr(fb.Y_c), C.byref(p), ptr(Sn),
               ptr(alive_s), ptr(fb.O_c), ptr(fb.flags), ptr(fb.relu_c), None, None, 3 if fold else 1, st)
        if fold:
            return
    else:
        _perceive_pl(grid, pl, g, xv4, S, p, Y_s, fb.flags, st)
        _timed("step_mlp", lib.sphnca_step_mlp_list, N, ptr(fb.rows), ptr(fb.count), ptr(S), ptr(Y_s), C.byref(p), ptr(Sn), ptr(alive_s),
               ptr(fb.O_c), ptr(fb.flags), ptr(fb.relu_c), ptr(fb.A_c), ptr(fb.Y_c), 0, st)
    _mask(use_alpha, g, N, xv4, alive_s, Sn, fb.flags, st)


def _draw(N, dev, fire_index):
    """One step's fire draw: torch.rand(N), or -- fire_index = (n_draw, index [N]) -- entries `index` of torch.rand(n_draw): the
    periodic planes draw for their ghost particles too (reference nca.py:100 inside :245-299), so the draws of the N real
    particles are a subset of a longer call; consuming the generator identically keeps seeded runs comparable."""
    if fire_index is None:
        return torch.rand(N, device=dev)
    n_draw, index = fire_index
    return torch.rand(n_draw, device=dev)[index]


class _UnpackRows(torch.autograd.Function):
    """cell-sorted rows -> the caller's particle order (adjoint: the inverse permutation)."""

    @staticmethod
    def forward(ctx, S, grid):
        ctx.grid = grid
        return grid.unpack_rows(S, S.shape[1])

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dA):
        return ctx.grid.pack_rows(require_cuda_f32(dA, "dA"), dA.shape[1]), None


class _NCARollout(torch.autograd.Function):
    """K chained NCA steps as ONE autograd node (replaces the K-node tape the
    reference builds in the loop of code/nca.py:141-145).  The state never leaves cell-sorted order:
    per step the forward is perceive -> MLP -> mask, the backward data-gradient -> weight-gradient ->
    sweep adjoint, with the parameter gradients accumulated in place over all steps (fixed order:
    deterministic).  Outputs: the K+1 states S_0..S_K in SORTED order (unpack lazily, `RolloutStates`);
    a loss may read any subset of them -- missing output gradients are skipped, not materialised.
    `taped`: 0 = no tape (inference), 1 = keep every step's MLP input / output rows, ReLU bits, fire list and flags
    (TAPE_BYTES = 373 B per particle-step besides the states), 2 = activation checkpointing: keep only the states and
    the packed fire bytes (65 B per particle-step) and re-run list -> perceive -> MLP -> mask of a step right before
    its backward; the kernels are deterministic, so both modes give bit-identical gradients.
    With `FIRE_LISTS` (default) the MLP half of a step -- and, without an alpha channel, its perception sweep -- runs
    over the particles whose fire draw came up only (csrc/firelist.cu); the others are carried over.
    The fire mask of step k comes from `torch.rand(N)`, the same call, in the same order, as the reference's
    `SPHNCA.forward` (nca.py:100), compared with the rate by the list kernel (`u <= fire_rate` in fp32, like torch)."""

    @staticmethod
    def forward(ctx, x, v, A0, W1, b1, W2, b2, grid, cfg, steps, fire_rate, keep_all, taped, groups=None, fire_index=None):
        if groups is not None:
            return _microbatched_forward(ctx, x, v, A0, W1, b1, W2, b2, grid, cfg, steps, fire_rate, keep_all, taped, groups, fire_index)
        ctx.groups = None
        update_rule, use_alpha, perception_scale, orig_rate = cfg
        x = require_cuda_f32(x, "x"); v = require_cuda_f32(v, "v"); A0 = require_cuda_f32(A0, "A")
        W1c, b1c, W2c, b2c = [require_cuda_f32(t, n) for t, n in ((W1, "W1"), (b1, "b1"), (W2, "W2"), (b2, "b2"))]
        N = x.shape[0]
        n_out = 33 if update_rule == "gated" else 16
        if A0.shape != (N, 16) or W1c.shape != (256, 48) or W2c.shape != (n_out, 256):
            raise ValueError("fused rollout needs A [N,16], W1 [256,48], W2 [33|16,256]")
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        ctx.set_materialize_grads(False)
        xv4 = grid.sorted_xv(x, v)
        p = _step_params(W1c, b1c, W2c, b2c, update_rule, use_alpha, perception_scale, orig_rate)
        g, st = grid.c_grid(), stream_ptr()
        S = grid.pack_rows(A0, 16)
        states, saved = [S], []
        alive_s = torch.empty((N,), **u8)
        Y_s = O_s = flags_s = relu_s = None
        fl = _fire_lists_on() and N > 0
        fbuf = ws_fl = None
        if fl:
            ws_fl = _fire_list_workspace(N, dev)
            Y_s = torch.empty((N, 32), **f32)                 # perception of the current step: scratch, the tape keeps Y_c
        fire_s = None
        alive_cur = None
        for _ in range(steps):
            u = _draw(N, dev, fire_index)                     # the reference's draw (nca.py:100), one call per step
            if u.dtype != torch.float32 or not u.is_contiguous():
                u = u.to(torch.float32).contiguous()
            Sn = torch.empty((N, 16), **f32)
            if fl:
                # MLP half over the fired rows only (firelist.cu); not-fired rows are carried over by the list kernel,
                # which also turns the draw into the (cell-sorted) fire bytes
                if taped == 1 or fbuf is None:
                    fbuf = _FireBuffers(N, dev, compact_inputs=True)
                if taped == 2 or fire_s is None:
                    fire_s = torch.empty((N,), **u8)
                alive_io = None
                if use_alpha:                                 # alive bytes of the state chained from step to step (mask list sweeps)
                    if alive_cur is None:
                        alive_cur = (S[:, 3] > 0.1).view(torch.uint8).contiguous()
                    alive_io = (alive_cur, torch.empty((N,), **u8))
                _step_forward_lists(grid, x, v, g, xv4, S, fire_s, p, use_alpha, Y_s, Sn, alive_s, fbuf, ws_fl, st,
                                    draw=(u, fire_rate), alive_io=alive_io)
                if use_alpha:
                    alive_cur = alive_io[1]
                if taped == 1:
                    saved += fbuf.tensors()
                    if TAPE_PROBE is not None:
                        TAPE_PROBE.append((fbuf.rows, fbuf.count, fbuf.relu_c, fbuf.Y_c, fbuf.O_c))
                elif taped == 2:
                    saved.append(fire_s)
                if keep_all or taped:
                    states.append(Sn)
                else:
                    states = [states[0], Sn]
                S = Sn
                continue
            fire_s = grid.pack_bytes((u <= fire_rate).view(torch.uint8))
            if taped == 1 or Y_s is None:
                Y_s = torch.empty((N, 32), **f32); O_s = torch.empty((N, 36), **f32)
                flags_s = torch.empty((N,), **u8); relu_s = torch.empty((N, 8), dtype=torch.int32, device=dev)
            if N:
                if SPLIT_FORWARD:
                    _perceive(grid, x, v, g, xv4, S, p, Y_s, flags_s, st)
                    _timed("step_mlp", lib.sphnca_step_mlp, N, ptr(S), ptr(Y_s), ptr(fire_s), C.byref(p), ptr(Sn),
                           ptr(alive_s), ptr(O_s), ptr(flags_s), ptr(relu_s), st)
                else:
                    _timed("step_fwd", lib.sphnca_step_fwd, g, ptr(xv4), ptr(S), ptr(fire_s), C.byref(p), ptr(Sn),
                           ptr(alive_s), ptr(Y_s), ptr(O_s), ptr(flags_s), ptr(relu_s), st)
                _mask(use_alpha, g, N, xv4, alive_s, Sn, flags_s, st)
            if taped == 1:
                saved += [Y_s, O_s, flags_s, relu_s]
            elif taped == 2:
                saved.append(fire_s)
            if keep_all or taped:
                states.append(Sn)
            else:
                states = [states[0], Sn] if len(states) else [Sn]
            S = Sn
        ctx.grid, ctx.cfg, ctx.steps, ctx.taped, ctx.fire_lists = grid, cfg, steps, taped, fl
        ctx.pairs = grid.pair_lists(x, v, blur=bool(use_alpha)) if (N and taped) else None
        if taped:
            ctx.save_for_backward(xv4, W1c, b1c, W2c, b2c, *states, *saved)
        return tuple(states)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *dstates):
        if ctx.groups is not None:
            return _microbatched_backward(ctx, dstates)
        K = ctx.steps
        t = ctx.saved_tensors
        xv4, W1, b1, W2, b2 = t[:5]
        states, saved = t[5:5 + K + 1], t[5 + K + 1:]
        grid = ctx.grid
        N = xv4.shape[0]
        dev = xv4.device
        f32 = dict(dtype=torch.float32, device=dev)
        need = ctx.needs_input_grad
        want_w = any(need[3:7])
        p = _step_params(W1, b1, W2, b2, *ctx.cfg)
        dW1 = db1 = dW2 = db2 = None
        ws, nb = None, C.c_size_t(0)
        if want_w:
            dW1, db1, dW2, db2 = torch.zeros_like(W1), torch.zeros_like(b1), torch.zeros_like(W2), torch.zeros_like(b2)
            check(lib.sphnca_mlp_wgrad_workspace_bytes(max(N, 1), C.byref(nb)))
            ws = torch.empty((nb.value,), dtype=torch.uint8, device=dev)
        dO_s = torch.empty((N, 36), **f32)
        dY_s = torch.empty((N, 32), **f32)
        dAdir_s = torch.empty((N, 16), **f32)
        own = [torch.empty((N, 16), **f32), torch.empty((N, 16), **f32)]   # gradient ping-pong (never write autograd's buffers)
        st = stream_ptr()
        recompute = ctx.taped == 2
        fl = ctx.fire_lists
        # fire lists + pair lists: the data-gradient kernel writes the dL/dY rows as [x quarter | y quarter] pairs, the layout the
        # list adjoint gathers with one 32-byte load per lane (include/sphnca.h: sphnca_step_params.dy_interleaved)
        p.dy_interleaved = int(bool(fl) and ctx.pairs is not None)
        # The weight-gradient kernel of step k (tensor-core bound) and its adjoint sweep (latency bound) both depend on the
        # data-gradient kernel only: with OVERLAP_WGRAD the weight-gradient kernels run on a side stream (serialised there, so the
        # parameter gradients accumulate in the same fixed order) beside the adjoint sweep and the next step's data gradient.
        # (Not while bench.py times the kernels one by one: PROFILE.)  Measured on configs[1]: 808 -> 830 M particle-steps/s; the
        # weight-gradient CTAs (416 threads x 128 registers, 223 KB shared memory) leave room for only 2 of the 10 adjoint CTAs an
        # SM normally holds, so the two kernels overlap by ~0.03 ms per step, not by the 0.15 ms the shorter one lasts.
        side = _side_stream(dev) if (OVERLAP_WGRAD and PROFILE is None and fl and want_w and not recompute and N) else None
        dO_alt = torch.empty((N, 36), **f32) if side is not None else None       # dO double buffer: step k-1 writes while wgrad(k) reads
        wg_done = [None, None]
        wg_calls = 0                                          # launches that added to the weight-gradient partials in `ws`
        if fl:
            g = grid.c_grid()
            dO_c = dO_s                                       # dL/d(MLP output) of the listed rows, by list position
            if recompute:
                fbuf = _FireBuffers(N, dev, compact_inputs=True)
                Y_s = torch.empty((N, 32), **f32)
                Sn_scr, alive_s = torch.empty((N, 16), **f32), torch.empty((N,), dtype=torch.uint8, device=dev)
                ws_fl = _fire_list_workspace(N, dev)
        elif recompute and N:
            u8 = dict(dtype=torch.uint8, device=dev)
            Y_s = torch.empty((N, 32), **f32); O_s = torch.empty((N, 36), **f32)
            flags_s = torch.empty((N,), **u8); relu_s = torch.empty((N, 8), dtype=torch.int32, device=dev)
            Sn_scr, alive_s = torch.empty((N, 16), **f32), torch.empty((N,), **u8)
            g = grid.c_grid()
        gcur = dstates[K]
        if gcur is not None:
            gcur = require_cuda_f32(gcur, "dS")
        for k in range(K - 1, -1, -1):
            if gcur is not None and N and fl:
                if recompute:                   # the forward of step k again, from the saved state and fire bytes
                    _step_forward_lists(grid, None, None, g, xv4, states[k], saved[k], p, ctx.cfg[1], Y_s, Sn_scr, alive_s, fbuf,
                                        ws_fl, st, pl=ctx.pairs)
                    fb = fbuf
                else:
                    fb = _FireBuffers.from_tensors(saved[_FireBuffers.PER_STEP * k:_FireBuffers.PER_STEP * (k + 1)])
                pl = ctx.pairs
                if pl is None:      # the cell-sweep adjoint reads every row: the not-fired ones are written first
                    _timed("step_bwd_unfired", lib.sphnca_step_bwd_unfired, N, ptr(fb.flags), ptr(gcur), ptr(dY_s), ptr(dAdir_s), st)
                if side is not None:
                    dO_c = dO_s if (k & 1) else dO_alt
                    if wg_done[k & 1] is not None:          # the weight-gradient kernel that last read this dO buffer
                        torch.cuda.current_stream(dev).wait_event(wg_done[k & 1])
                _timed("step_bwd", lib.sphnca_step_bwd_list, N, ptr(fb.rows), ptr(fb.count), ptr(fb.A_c), ptr(fb.O_c), ptr(fb.flags),
                       ptr(fb.relu_c), ptr(gcur), C.byref(p), ptr(dO_c), ptr(dY_s), ptr(dAdir_s), st)
                if want_w and side is not None:
                    ready = torch.cuda.Event(); ready.record(torch.cuda.current_stream(dev))
                    side.wait_event(ready)
                    # per-stream partial sums accumulate over the steps; ONE reduction into the gradients after the loop
                    check(lib.sphnca_mlp_wgrad_list_partial(N, ptr(fb.count), ptr(fb.A_c), ptr(fb.Y_c), ptr(dO_c), C.byref(p), wg_calls > 0,
                                                            ptr(ws), nb.value, C.c_void_p(side.cuda_stream)))
                    wg_calls += 1
                    wg_done[k & 1] = torch.cuda.Event(); wg_done[k & 1].record(side)
                elif want_w:
                    _timed("mlp_wgrad", lib.sphnca_mlp_wgrad_list_partial, N, ptr(fb.count), ptr(fb.A_c), ptr(fb.Y_c), ptr(dO_c), C.byref(p),
                           wg_calls > 0, ptr(ws), nb.value, st)
                    wg_calls += 1
                out = own[k & 1]
                if pl is not None:  # fire-aware list adjoint: rows of not-fired particles come from their flag byte
                    _timed("step_adjoint", lib.sphnca_step_adjoint_pairs_fired, grid.c_grid(), ptr(pl["slots"]), ptr(pl["base"]), ptr(pl["J"]),
                           ptr(pl["WA"]), ptr(pl["SXY"]), ptr(xv4), ptr(dY_s), ptr(dAdir_s), ptr(fb.flags), ptr(gcur), C.byref(p), ptr(out), st)
                else:
                    _adjoint(grid, None, xv4, dY_s, dAdir_s, p, out, st)
                gcur = out
            elif gcur is not None and N:        # steps after the last state a loss reads carry no gradient
                if recompute:                   # the forward of step k again, from the saved state and fire bytes
                    _perceive_pl(grid, ctx.pairs, g, xv4, states[k], p, Y_s, flags_s, st)
                    _timed("step_mlp", lib.sphnca_step_mlp, N, ptr(states[k]), ptr(Y_s), ptr(saved[k]), C.byref(p),
                           ptr(Sn_scr), ptr(alive_s), ptr(O_s), ptr(flags_s), ptr(relu_s), st)
                    _mask(ctx.cfg[1], g, N, xv4, alive_s, Sn_scr, flags_s, st)
                else:
                    Y_s, O_s, flags_s, relu_s = saved[4 * k:4 * k + 4]
                _timed("step_bwd", lib.sphnca_step_bwd, N, ptr(states[k]), ptr(Y_s), ptr(O_s), ptr(flags_s), ptr(relu_s),
                       ptr(gcur), C.byref(p), ptr(dO_s), ptr(dY_s), ptr(dAdir_s), None, None, st)
                if want_w:
                    _timed("mlp_wgrad", lib.sphnca_mlp_wgrad, N, ptr(states[k]), ptr(Y_s), ptr(dO_s), None, None,
                           C.byref(p), ptr(dW1), ptr(db1), ptr(dW2), ptr(db2), ptr(ws), nb.value, st)
                out = own[k & 1]
                _adjoint(grid, ctx.pairs, xv4, dY_s, dAdir_s, p, out, st)
                gcur = out
            if dstates[k] is not None:
                dk = require_cuda_f32(dstates[k], "dS")
                gcur = dk if gcur is None else gcur.add_(dk)
        if wg_calls:
            check(lib.sphnca_mlp_wgrad_reduce(N, C.byref(p), ptr(dW1), ptr(db1), ptr(dW2), ptr(db2), ptr(ws), nb.value,
                                              C.c_void_p(side.cuda_stream) if side is not None else st))
        if side is not None:
            torch.cuda.current_stream(dev).wait_stream(side)
        dA0 = None
        if need[2]:
            dA0 = grid.unpack_rows(gcur, 16) if gcur is not None else torch.zeros((N, 16), **f32)
        return (None, None, dA0, dW1 if need[3] else None, db1 if need[4] else None, dW2 if need[5] else None,
                db2 if need[6] else None, None, None, None, None, None, None, None, None)


# ---- micro-batches over sections ------------------------------------------------------------------------------------
# Batch elements (`sections`) are independent through the whole rollout (reference operators_batch.py:34-42: separate
# OFFSET / CELL_OFFSET ranges) and contiguous in the cell-sorted arrays.  When the pair lists (~0.8-1.1 KB per
# particle) and the per-step work buffers of the WHOLE batch would not fit next to the K+1 saved states -- BASELINE
# configs[4]: 64 clouds x 1,000,000 particles -- the node runs GROUPS of consecutive sections one after the other, group
# outermost and step innermost: a group's pair lists and work buffers live only while that group runs (forward: built,
# K steps, dropped; backward: rebuilt once, K steps back), the K+1 states and K fire-byte arrays of all particles are the only
# O(N K) storage (65 B per particle-step: this mode always checkpoints), parameter gradients accumulate over groups in a
# fixed order.  States and dL/dA0 are bit-identical to the whole-batch node (row-independent arithmetic, same lists); the
# fire draws are the reference's: one torch.rand(N) per step, in step order (nca.py:100), drawn before the first group runs.
MICROBATCH = "auto"            # "auto" | None (never) | int (sections per group; tests force small groups)
MICROBATCH_WORK_BYTES = 373 + 128 + 16 + 144 + 128 + 64 + 128 + 64     # fire buffers, Y, Sn scratch, dO, dY, dAdir, gradient ping-pong, slack
MICROBATCH_FRACTION = 0.20      # of the device memory torch does not hold and the states will not take (the loss side of a
                                # training iteration unpacks the states it reads and holds their gradients: ~10 x 64 B per particle)


def plan_microbatches(grid, steps, use_alpha, device, taped, setting=None):
    """None = run the whole batch at once, else [(first section, last + 1)] groups of consecutive sections."""
    setting = MICROBATCH if setting is None else setting
    B, N = len(grid.sections), grid.num_particles
    if setting is None or B <= 1 or N == 0 or not _fire_lists_on() or not _lib.USE_PAIR_LISTS or _lib.FORCE_SWEEP_PATH is not None:
        return None
    if isinstance(setting, int) and not isinstance(setting, bool):
        per = max(1, setting)
        return [(s, min(s + per, B)) for s in range(0, B, per)] if per < B else None
    per_particle = grid.PAIR_SLOTS_ESTIMATE * grid.PAIR_ENTRY_BYTES[bool(use_alpha)] + MICROBATCH_WORK_BYTES
    free = torch.cuda.get_device_properties(device).total_memory - torch.cuda.memory_allocated(device)
    free -= N * (64 * (steps + 1) + steps) if taped else N * 128
    budget = MICROBATCH_FRACTION * max(free, 0)
    if N * per_particle <= budget:
        return None
    groups, s0, acc = [], 0, 0
    for s, n in enumerate(grid.sections):
        if s > s0 and (acc + n) * per_particle > budget:
            groups.append((s0, s)); s0, acc = s, 0
        acc += n
    groups.append((s0, B))
    return groups if len(groups) > 1 else None


class _GroupWork:
    """Work buffers of one micro-batch, sized for the largest group and reused by all groups."""

    def __init__(self, n_max, dev, backward, want_w):
        f32 = dict(dtype=torch.float32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        self.fb = _FireBuffers(n_max, dev, compact_inputs=True)
        self.Y_s = torch.empty((n_max, 32), **f32)
        self.alive_s = torch.empty((n_max,), **u8)
        self.ws_fl = _fire_list_workspace(n_max, dev)
        self.ping = [torch.empty((n_max, 16), **f32), torch.empty((n_max, 16), **f32)]   # state (inference) / gradient ping-pong
        if backward:
            self.dO_c = torch.empty((n_max, 36), **f32)
            self.dY_s = torch.empty((n_max, 32), **f32)
            self.dAdir_s = torch.empty((n_max, 16), **f32)
            self.Sn_scr = torch.empty((n_max, 16), **f32)
            self.nb = C.c_size_t(0)
            self.ws = None
            if want_w:
                check(lib.sphnca_mlp_wgrad_workspace_bytes(max(n_max, 1), C.byref(self.nb)))
                self.ws = torch.empty((self.nb.value,), **u8)


def _group_lists(sub, xv, use_alpha):
    # a group was sized to fit: build its lists unconditionally (the budget check of Hashgrid.pair_lists is for whole batches)
    return sub.build_pair_lists(xv, blur=bool(use_alpha), budget_bytes=float("inf")) if sub.num_particles else None


def _microbatched_forward(ctx, x, v, A0, W1, b1, W2, b2, grid, cfg, steps, fire_rate, keep_all, taped, groups, fire_index=None):
    update_rule, use_alpha, perception_scale, orig_rate = cfg
    x = require_cuda_f32(x, "x"); v = require_cuda_f32(v, "v"); A0 = require_cuda_f32(A0, "A")
    W1c, b1c, W2c, b2c = [require_cuda_f32(t, n) for t, n in ((W1, "W1"), (b1, "b1"), (W2, "W2"), (b2, "b2"))]
    N = x.shape[0]
    n_out = 33 if update_rule == "gated" else 16
    if A0.shape != (N, 16) or W1c.shape != (256, 48) or W2c.shape != (n_out, 256):
        raise ValueError("fused rollout needs A [N,16], W1 [256,48], W2 [33|16,256]")
    dev = x.device
    ctx.set_materialize_grads(False)
    xv4 = grid.sorted_xv(x, v)
    p = _step_params(W1c, b1c, W2c, b2c, update_rule, use_alpha, perception_scale, orig_rate)
    st = stream_ptr()
    S0 = grid.pack_rows(A0, 16)
    # the reference's draws, one torch.rand(N) per step in step order, kept as bytes in PARTICLE order (N bytes per step)
    fire_po = [(_draw(N, dev, fire_index) <= fire_rate).view(torch.uint8) for _ in range(steps)]
    every = bool(keep_all or taped)
    f32 = dict(dtype=torch.float32, device=dev)
    states = [S0] + [torch.empty((N, 16), **f32) for _ in range(steps if every else min(steps, 1))]
    subs = [grid.subgrid(s0, s1) for s0, s1 in groups]
    work = _GroupWork(max(s.num_particles for s in subs), dev, backward=False, want_w=False)
    for sub in subs:
        p0, p1 = sub.range
        n = p1 - p0
        if n == 0:
            continue
        g, xv = sub.c_grid(), xv4[p0:p1]
        pl = _group_lists(sub, xv, use_alpha)
        S = S0[p0:p1]
        for k in range(steps):
            fire_s = sub.pack_bytes(fire_po[k])
            if every:
                Sn = states[k + 1][p0:p1]
            else:                                   # inference: only the final state is kept
                Sn = states[-1][p0:p1] if k == steps - 1 else work.ping[k & 1][:n]
            _step_forward_lists(sub, None, None, g, xv, S, fire_s, p, use_alpha, work.Y_s, Sn, work.alive_s, work.fb, work.ws_fl, st, pl=pl)
            S = Sn
        del pl
    ctx.grid, ctx.cfg, ctx.steps, ctx.taped, ctx.fire_lists, ctx.groups = grid, cfg, steps, 2 if taped else 0, True, groups
    ctx.pairs = None
    if taped:
        ctx.save_for_backward(xv4, W1c, b1c, W2c, b2c, *states, *fire_po)
    return tuple(states)


def _microbatched_backward(ctx, dstates):
    K = ctx.steps
    t = ctx.saved_tensors
    xv4, W1, b1, W2, b2 = t[:5]
    states, fire_po = t[5:5 + K + 1], t[5 + K + 1:]
    grid, use_alpha = ctx.grid, ctx.cfg[1]
    N, dev = xv4.shape[0], xv4.device
    need = ctx.needs_input_grad
    want_w = any(need[3:7])
    p = _step_params(W1, b1, W2, b2, *ctx.cfg)
    dW1 = db1 = dW2 = db2 = None
    if want_w:
        dW1, db1, dW2, db2 = torch.zeros_like(W1), torch.zeros_like(b1), torch.zeros_like(W2), torch.zeros_like(b2)
    st = stream_ptr()
    subs = [grid.subgrid(s0, s1) for s0, s1 in ctx.groups]
    work = _GroupWork(max(s.num_particles for s in subs), dev, backward=True, want_w=want_w)
    dstates = [None if d is None else require_cuda_f32(d, "dS") for d in dstates]
    dA0_s = torch.zeros((N, 16), dtype=torch.float32, device=dev) if need[2] else None
    if any(d is not None for d in dstates):
        for sub in subs:
            p0, p1 = sub.range
            n = p1 - p0
            if n == 0:
                continue
            g, xv = sub.c_grid(), xv4[p0:p1]
            pl = _group_lists(sub, xv, use_alpha)
            p.dy_interleaved = int(pl is not None)            # see _NCARollout.backward
            fb = work.fb
            gcur = dstates[K][p0:p1] if dstates[K] is not None else None
            wg_calls = 0                                      # a group's steps add to ITS partials (the stream count follows n)
            for k in range(K - 1, -1, -1):
                if gcur is not None:
                    fire_s = sub.pack_bytes(fire_po[k])
                    # the forward of step k again, from the saved state and fire bytes (activation checkpointing)
                    _step_forward_lists(sub, None, None, g, xv, states[k][p0:p1], fire_s, p, use_alpha, work.Y_s, work.Sn_scr, work.alive_s,
                                        fb, work.ws_fl, st, pl=pl)
                    if pl is None:
                        _timed("step_bwd_unfired", lib.sphnca_step_bwd_unfired, n, ptr(fb.flags), ptr(gcur), ptr(work.dY_s), ptr(work.dAdir_s), st)
                    _timed("step_bwd", lib.sphnca_step_bwd_list, n, ptr(fb.rows), ptr(fb.count), ptr(fb.A_c), ptr(fb.O_c), ptr(fb.flags),
                           ptr(fb.relu_c), ptr(gcur), C.byref(p), ptr(work.dO_c), ptr(work.dY_s), ptr(work.dAdir_s), st)
                    if want_w:
                        _timed("mlp_wgrad", lib.sphnca_mlp_wgrad_list_partial, n, ptr(fb.count), ptr(fb.A_c), ptr(fb.Y_c), ptr(work.dO_c),
                               C.byref(p), wg_calls > 0, ptr(work.ws), work.nb.value, st)
                        wg_calls += 1
                    out = work.ping[k & 1][:n]
                    if pl is not None:
                        _timed("step_adjoint", lib.sphnca_step_adjoint_pairs_fired, g, ptr(pl["slots"]), ptr(pl["base"]), ptr(pl["J"]),
                               ptr(pl["WA"]), ptr(pl["SXY"]), ptr(xv), ptr(work.dY_s), ptr(work.dAdir_s), ptr(fb.flags), ptr(gcur), C.byref(p),
                               ptr(out), st)
                    else:
                        _adjoint(sub, None, xv, work.dY_s, work.dAdir_s, p, out, st)
                    gcur = out
                if dstates[k] is not None:
                    dk = dstates[k][p0:p1]
                    gcur = dk if gcur is None else gcur.add_(dk)
            if wg_calls:
                check(lib.sphnca_mlp_wgrad_reduce(n, C.byref(p), ptr(dW1), ptr(db1), ptr(dW2), ptr(db2), ptr(work.ws), work.nb.value, st))
            if dA0_s is not None and gcur is not None:
                dA0_s[p0:p1].copy_(gcur)
            del pl
    dA0 = grid.unpack_rows(dA0_s, 16) if need[2] else None
    return (None, None, dA0, dW1 if need[3] else None, db1 if need[4] else None, dW2 if need[5] else None,
            db2 if need[6] else None, None, None, None, None, None, None, None, None)


class RolloutStates:
    """The `out` list of the reference's rollout drivers (nca.py:140-146: one `A.clone()` per step) as a lazy
    sequence: states live cell-sorted on the device and item k is permuted back to particle order -- through
    autograd -- only when somebody reads it (the loss reads 5 of K+1)."""

    def __init__(self, sorted_states, grid):
        self._s, self._grid, self._cache = list(sorted_states), grid, {}

    def __len__(self):
        return len(self._s)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self._s)))]
        if k < 0:
            k += len(self._s)
        if not 0 <= k < len(self._s):
            raise IndexError(k)
        if k not in self._cache:
            self._cache[k] = _UnpackRows.apply(self._s[k], self._grid)
        return self._cache[k]

    def __iter__(self):
        return (self[k] for k in range(len(self._s)))


# Run the weight-gradient kernels of a taped backward on a side stream, beside the adjoint sweeps (see _NCARollout.backward).
OVERLAP_WGRAD = __import__("os").environ.get("SPHNCA_OVERLAP_WGRAD", "1") != "0"
_SIDE_STREAMS = {}


def _side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


# Activation checkpointing inside the rollout node: True / False, or "auto" = checkpoint when
# the full tape (TAPE_BYTES per particle-step) would take more than AUTO_CHECKPOINT_FRACTION of the HBM torch does not hold.
CHECKPOINT = "auto"
TAPE_BYTES = 4 * (32 + 36 + 8) + 1 + 4 * (16 + 1)       # + compact state rows and the list on the fire-list path
AUTO_CHECKPOINT_FRACTION = 0.5


def _tape_mode(N, steps, device, checkpoint):
    if checkpoint is None:
        checkpoint = CHECKPOINT
    if checkpoint == "auto":
        # the device's capacity minus what torch holds live (host-side counters: cudaMemGetInfo costs ~25 ms per call
        # on a busy device and this runs once per rollout; torch's cached-but-free blocks are reusable)
        free = torch.cuda.get_device_properties(device).total_memory - torch.cuda.memory_allocated(device)
        checkpoint = N * steps * TAPE_BYTES > AUTO_CHECKPOINT_FRACTION * free
    return 2 if checkpoint else 1


def nca_rollout(x, v, A, W1, b1, W2, b2, h, grid, steps, fire_rate, update_rule="gated", use_alpha=True,
                perception_scale=1.0, orig_rate=1.0, keep_all=True, checkpoint=None, microbatch=None, fire_index=None):
    """`steps` NCA steps from state A [N,16] as one autograd node -> RolloutStates (K+1 states if `keep_all`
    or anything requires grad, else [first, last]).  `checkpoint`: None = module default `CHECKPOINT`;
    `microbatch`: None = module default `MICROBATCH` (see plan_microbatches); `fire_index`: see _draw."""
    if "sphnca_step_fwd" in _lib.MISSING_SYMBOLS:
        raise RuntimeError("libsphnca.so was built without the fused step (step.cu)")
    hval = h.item() if isinstance(h, torch.Tensor) else float(h)
    if abs(hval - grid.h) > 1e-9 * abs(grid.h):
        raise ValueError(f"h={hval} differs from the grid's cell size {grid.h}")
    taped = int(torch.is_grad_enabled() and any(t.requires_grad for t in (A, W1, b1, W2, b2)))
    if taped:
        taped = _tape_mode(x.shape[0], int(steps), x.device, checkpoint)
    groups = plan_microbatches(grid, int(steps), use_alpha, x.device, taped, microbatch)
    states = _NCARollout.apply(x, v, A, W1, b1, W2, b2, grid, (update_rule, use_alpha, perception_scale, orig_rate),
                               int(steps), fire_rate, keep_all, taped, groups, fire_index)
    return RolloutStates(states, grid)


# True (default): call the two halves of the split forward (perception sweep, tcgen05 MLP) as separate
# C-ABI entries; False: one sphnca_step_fwd call, whose implementation follows sphnca_set_mlp_path.
SPLIT_FORWARD = True
# True: CUDA-core backward (step_bwd_kernel streams transposed hidden activations to mlp_wgrad_kernel);
# False (default): tensor-core backward (step_bwd_tc_kernel + mlp_wgrad_tc_kernel, no scratch).
SIMT_BACKWARD = False
# particles per backward chunk on the CUDA-core path (bounds the 2 KB/particle scratch to 8 GB)
WGRAD_CHUNK = 4 * 1024 * 1024


def frame_step_available(grid, x, v):
    """The tangent-frame step needs the pair lists (with the z weights); False when they are off or do not fit."""
    return "sphnca_step_perceive_pairs_frame" not in _lib.MISSING_SYMBOLS and x.shape[0] > 0 \
        and grid.pair_lists(x, v, blur=True, z=True) is not None


def nca_step(x, v, A, update_mask, W1, b1, W2, b2, h, grid, update_rule="gated", use_alpha=True,
             perception_scale=1.0, orig_rate=1.0, frame=None):
    """One NCA step.  x [N,3], v [N], A [N,16], update_mask bool [N] (the fire draw) -> new A [N,16].
    frame [N,8] = {T, 0, B, 0} per particle: the perception gradients are taken along T and B instead of x and y
    (surface clouds, `nca.feature_process_tangent`); forward only."""
    if "sphnca_step_fwd" in _lib.MISSING_SYMBOLS:
        raise RuntimeError("libsphnca.so was built without the fused step (step.cu)")
    hval = h.item() if isinstance(h, torch.Tensor) else float(h)
    if abs(hval - grid.h) > 1e-9 * abs(grid.h):
        raise ValueError(f"h={hval} differs from the grid's cell size {grid.h}")
    if frame is not None:
        frame = require_cuda_f32(frame, "frame")
        with torch.no_grad():
            return _NCAStep.apply(x, v, A, update_mask, W1, b1, W2, b2, h, grid, update_rule, use_alpha,
                                  perception_scale, orig_rate, frame)
    return _NCAStep.apply(x, v, A, update_mask, W1, b1, W2, b2, h, grid, update_rule, use_alpha,
                          perception_scale, orig_rate)
