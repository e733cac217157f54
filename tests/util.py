"""Shared helpers for the parity tests."""
import contextlib
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# fp32 parity bar of BASELINE.json's north_star: |new - ref| <= 1e-5 * max|ref| per tensor
RTOL = 1e-5


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel_err(a, ref):
    a = np.asarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    ref = np.asarray(ref.detach().cpu().numpy() if isinstance(ref, torch.Tensor) else ref, dtype=np.float64)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30))


def assert_close(a, ref, rtol=RTOL, what=""):
    e = rel_err(a, ref)
    assert e <= rtol, f"{what}: rel err {e:.3e} > {rtol:g}"
    return e


def jittered_lattice(n_side, dim, gen, jitter=1.0):
    """(i+0.5)*s - 1 in [-1,1]^dim + (U-0.5)*jitter*s  -- the synthetic inputs of SURVEY.md 8d."""
    s = 2.0 / n_side
    axes = [(torch.arange(n_side, dtype=torch.float32) + 0.5) * s - 1.0 for _ in range(dim)]
    mesh = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(-1, dim)
    return mesh + (torch.rand(mesh.shape, generator=gen) - 0.5) * jitter * s


@contextlib.contextmanager
def forced_fire_masks(masks):
    """Make the model's `torch.rand(N, device=...) <= fire_rate` draws reproduce the given bool masks
    (the golden vectors recorded the CPU draws of the reference run)."""
    masks = list(masks)
    orig = torch.rand

    def fake(*size, **kw):
        m = masks.pop(0)
        m = torch.as_tensor(m)
        out = torch.where(m, torch.zeros(m.shape), torch.full(m.shape, 2.0))
        return out.to(kw.get("device", "cpu"))

    torch.rand = fake
    try:
        yield
    finally:
        torch.rand = orig


def check_tables_against(ref_idx, ref_cs, ref_ce, ref_np, grid, sections, total_cells):
    """cell_start / cell_end / num_points bit-exact; idx as per-cell index sets."""
    cs = grid.cell_start.cpu().numpy(); ce = grid.cell_end.cpu().numpy()
    idx = grid.idx.cpu().numpy()
    assert np.array_equal(grid.num_points.cpu().numpy(), ref_np)
    assert np.array_equal(cs, ref_cs)
    assert np.array_equal(ce, ref_ce)
    assert idx.dtype == np.int64
    off = 0
    for b, n in enumerate(sections):
        a = idx[off:off + n]; r = ref_idx[off:off + n]
        s = cs[b * total_cells:(b + 1) * total_cells]; e = ce[b * total_cells:(b + 1) * total_cells]
        occ = np.flatnonzero(s >= 0)
        # same members per cell <=> sorting each cell's slice gives equal arrays
        for c in occ:
            assert np.array_equal(np.sort(a[s[c]:e[c]]), np.sort(r[s[c]:e[c]])), f"cell {c} of section {b}"
        off += n
