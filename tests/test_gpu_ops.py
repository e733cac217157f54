"""GPU parity tests of the hash-grid build and the five sphops operators, called through the
reference-shaped Python interface -> C ABI -> sm_100a kernels.

Checked against (1) the committed golden vectors produced by the unmodified reference and
(2) the CPU oracle on larger seeded inputs, plus size-independent properties at BASELINE sizes.
Bar: integer tables bit-exact; floating point |new-ref| <= 1e-5 * max|ref| (tests/util.RTOL)."""
import numpy as np
import pytest
import torch

from tests.util import RTOL, assert_close, check_tables_against, jittered_lattice, load_golden, rel_err

pytestmark = pytest.mark.gpu
OPS_CASES = ["ops_3d_ragged", "ops_2d", "ops_padded2d", "ops_thin_axis", "ops_single_layer"]
DEV = "cuda"


def dims_arg(g):
    d = g["dims"].tolist()
    return tuple(int(v) for v in d)


@pytest.mark.parametrize("name", OPS_CASES)
def test_hashgrid_matches_reference_tables(name):
    import sphops
    g = load_golden(name)
    x = torch.from_numpy(g["x"]).to(DEV)
    grid = sphops.initialize_hashgrid(x, tuple(g["sections"]), float(g["h"]), dims_arg(g))
    check_tables_against(g["idx"], g["cell_start"], g["cell_end"], g["num_points"], grid,
                         g["sections"], int(g["total_cells"]))
    assert grid.total_cells == int(g["total_cells"])
    assert np.array_equal(grid.dims.cpu().numpy(), g["grid_dims"])
    # idx equals a STABLE sort by cell hash (what torch.sort gives the reference on CUDA)
    from oracle import sphnca_oracle as O
    og = O.build_hashgrid(torch.from_numpy(g["x"]), tuple(g["sections"]), float(g["h"]), g["dims"])
    assert np.array_equal(grid.idx.cpu().numpy(), og.idx)
    # lazily derived fields of the reference's dataclass
    for b, hs in enumerate(og.hi_sorted):
        occ, cnt = np.unique(hs, return_counts=True)
        assert np.array_equal(grid.occupied_cells[b].cpu().numpy(), occ)
        assert np.array_equal(grid.points_per_cells[b].cpu().numpy(), cnt)


@pytest.mark.parametrize("name", OPS_CASES)
def test_operators_match_reference(name):
    import sphops
    g = load_golden(name)
    x = torch.from_numpy(g["x"]).to(DEV)
    h = float(g["h"])
    grid = sphops.initialize_hashgrid(x, tuple(g["sections"]), h, dims_arg(g))
    xg = x.clone().requires_grad_(True)
    v = sphops.volume(xg, h, grid)
    assert_close(v, g["v"], what="volume")
    v.backward(torch.from_numpy(g["dLdv"]).to(DEV))
    assert_close(xg.grad, g["volume_dLdx"], what="volume dL/dx")
    v = v.detach()
    assert np.array_equal(sphops.count(x, h, grid).cpu().numpy(), g["count"])

    vref = torch.from_numpy(g["v"]).to(DEV)
    A = torch.from_numpy(g["A"]).to(DEV).requires_grad_(True)
    GA = sphops.gradient(x, vref, A, h, grid)
    assert_close(GA, g["GA"], what="gradient")
    GA.backward(torch.from_numpy(g["dLdGA"]).to(DEV))
    assert_close(A.grad, g["gradient_dLdA"], what="gradient dL/dA")
    for key in g:
        if key.startswith("blur_A"):
            F = key[len("blur_A"):]
            SA = sphops.blur(x, vref, torch.from_numpy(g[key]).to(DEV), h, grid)
            assert_close(SA, g["blur_SA" + F], what=f"blur F={F}")
    DA = sphops.divergence(x, vref, torch.from_numpy(g["div_A"]).to(DEV), h, grid)
    assert_close(DA, g["div_DA"], what="divergence")


def make_cloud(kind, seed):
    gen = torch.Generator().manual_seed(seed)
    if kind == "3d":
        x = torch.cat([jittered_lattice(22, 3, gen), jittered_lattice(18, 3, gen)])
        return x, (22 ** 3, 18 ** 3), 1.97 * 2 / 22, None
    if kind == "padded2d":
        x = torch.nn.functional.pad(jittered_lattice(96, 2, gen), (0, 1))
        return x, (96 * 96,), 3.2 * 2 / 96, None
    if kind == "2d":
        x = jittered_lattice(80, 2, gen)
        return x, (3000, 80 * 80 - 3000), 3.2 * 2 / 80, None
    raise KeyError(kind)


@pytest.mark.parametrize("kind", ["3d", "padded2d", "2d"])
def test_operators_match_oracle_midsize(kind):
    """~10-17k particles at the BASELINE neighbour counts (cand ~206/92, nbr ~31), oracle in fp64."""
    import math
    import sphops
    from oracle import sphnca_oracle as O
    x, sections, h, _ = make_cloud(kind, 11)
    D = x.shape[1]
    dims = math.ceil(2 / h)
    N = x.shape[0]
    gen = torch.Generator().manual_seed(5)
    og, pairs, v_ref = O.prepare(x, sections, h, dims)
    xd = x.to(DEV)
    grid = sphops.initialize_hashgrid(xd, sections, h, dims)
    check_tables_against(og.idx, og.cell_start, og.cell_end, og.num_points, grid, sections, og.total_cells)
    assert np.array_equal(grid.idx.cpu().numpy(), og.idx)
    v = sphops.volume(xd, h, grid)
    assert_close(v, v_ref, what="volume")
    assert np.array_equal(sphops.count(xd, h, grid).cpu().numpy(), O.count(pairs).numpy().astype(np.float32))

    for F in (16, 5, 1):
        A = torch.rand(N, F, generator=gen)
        G = torch.randn(N, F, D, generator=gen)
        Ad = A.to(DEV).requires_grad_(True)
        GA = sphops.gradient(xd, v, Ad, h, grid)
        assert_close(GA, O.gradient(pairs, v_ref, A), what=f"gradient F={F}")
        GA.backward(G.to(DEV))
        assert_close(Ad.grad, O.gradient_backward_dLdA_reference(pairs, v_ref, G), what=f"gradient dL/dA F={F}")

        Ab = A.double().requires_grad_(True)
        Gb = torch.randn(N, F, generator=gen)
        O.blur(pairs, v_ref, Ab).backward(Gb.double())
        Ad = A.to(DEV).requires_grad_(True)
        SA = sphops.blur(xd, v, Ad, h, grid)
        assert_close(SA, O.blur(pairs, v_ref, A), what=f"blur F={F}")
        SA.backward(Gb.to(DEV))
        assert_close(Ad.grad, Ab.grad, what=f"blur dL/dA (true adjoint) F={F}")

    for F in (3, 9):
        Av = torch.randn(N, F, D, generator=gen)
        Gd = torch.randn(N, F, generator=gen)
        Avd = Av.to(DEV).requires_grad_(True)
        DA = sphops.divergence(xd, v, Avd, h, grid)
        assert_close(DA, O.divergence(pairs, v_ref, Av), what=f"divergence F={F}")
        DA.backward(Gd.to(DEV))
        Avo = Av.double().requires_grad_(True)
        O.divergence(pairs, v_ref, Avo).backward(Gd.double())
        assert_close(Avd.grad, Avo.grad, what=f"divergence dL/dA (true adjoint) F={F}")

    dLdv = torch.rand(N, generator=gen)
    xg = xd.clone().requires_grad_(True)
    sphops.volume(xg, h, grid).backward(dLdv.to(DEV))
    assert_close(xg.grad, O.volume_backward_dLdx_reference(pairs, v_ref, dLdv), what="volume dL/dx")


def test_edge_cases():
    """Empty batch, empty sections, a lone particle, everything in one cell, h as a 0-d tensor."""
    import sphops
    from oracle import sphnca_oracle as O
    # empty packed batch
    x0 = torch.zeros((0, 3), device=DEV)
    g0 = sphops.initialize_hashgrid(x0, (), 0.5, 4)
    assert g0.idx.numel() == 0 and g0.cell_start.numel() == 0
    g1 = sphops.initialize_hashgrid(x0, (0, 0), 0.5, 4)
    assert g1.cell_start.numel() == 2 * 64 and int((g1.cell_start != -1).sum()) == 0
    assert sphops.volume(x0, 0.5, g1).shape == (0,)
    assert sphops.gradient(x0, torch.zeros(0, device=DEV), torch.zeros(0, 16, device=DEV), 0.5, g1).shape == (0, 16, 3)
    # ragged with an empty section in the middle, one lone particle, h passed as a tensor
    gen = torch.Generator().manual_seed(3)
    x = torch.rand(41, 3, generator=gen) * 2 - 1
    sections = (1, 0, 40)
    h = torch.tensor(0.5)
    og, pairs, v_ref = O.prepare(x, sections, 0.5, 4)
    xd = x.to(DEV)
    grid = sphops.initialize_hashgrid(xd, sections, h, torch.tensor([4, 4, 4], device=DEV))
    check_tables_against(og.idx, og.cell_start, og.cell_end, og.num_points, grid, sections, 64)
    v = sphops.volume(xd, h, grid)
    assert_close(v, v_ref, what="volume")
    A = torch.rand(41, 16, generator=gen)
    assert_close(sphops.gradient(xd, v, A.to(DEV), h, grid), O.gradient(pairs, v_ref, A), what="gradient")
    assert float(sphops.gradient(xd, v, A.to(DEV), h, grid)[0].abs().max()) == 0.0   # lone particle: no neighbours
    # all particles in ONE cell (maximum collision) incl. exact duplicates (r == 0 -> weight 0)
    x = torch.rand(300, 3, generator=gen) * 0.2 + 0.1
    x[10] = x[11]
    og, pairs, v_ref = O.prepare(x, (300,), 0.5, 4)
    xd = x.to(DEV)
    grid = sphops.initialize_hashgrid(xd, (300,), 0.5, 4)
    assert int((grid.cell_start >= 0).sum()) == 1
    assert np.array_equal(grid.idx.cpu().numpy(), np.arange(300))
    v = sphops.volume(xd, 0.5, grid)
    assert_close(v, v_ref, what="volume one-cell")
    A = torch.rand(300, 16, generator=gen)
    assert_close(sphops.gradient(xd, v, A.to(DEV), 0.5, grid), O.gradient(pairs, v_ref, A), what="gradient one-cell")
    assert np.array_equal(sphops.count(xd, 0.5, grid).cpu().numpy(), O.count(pairs).numpy().astype(np.float32))
    # wrong h for this grid is refused
    with pytest.raises(ValueError):
        sphops.volume(xd, 0.25, grid)


def test_properties_at_baseline_size():
    """BASELINE 3-D config size (100^3 = 1M particles): properties that need no oracle.
    (1) sortedness/permutation of the tables, (2) <G, grad(A)> == <grad^T(G), A> (adjointness, fp64 dots),
    (3) linearity and constant-field annihilation of the gradient, (4) blur of v-weighted ones ~ partition."""
    import math
    import sphops
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(0)
    n_side = 100
    x = jittered_lattice(n_side, 3, gen).to(DEV)
    N = x.shape[0]
    h = 1.97 * 2 / n_side
    dims = math.ceil(2 / h)
    grid = sphops.initialize_hashgrid(x, (N,), h, dims)
    idx = grid.idx
    assert torch.equal(torch.sort(idx).values, torch.arange(N, device=DEV))
    cell = grid.sorted_cell.long()
    assert bool((cell[1:] >= cell[:-1]).all())
    same = cell[1:] == cell[:-1]
    assert bool((idx[1:][same] > idx[:-1][same]).all())            # ascending inside a cell (stable)
    cs, ce = grid.cell_start.long(), grid.cell_end.long()
    occ = cs >= 0
    assert int((ce[occ] - cs[occ]).sum()) == N
    v = sphops.volume(x, h, grid)
    cnt = sphops.count(x, h, grid)
    assert 25 < float(cnt.mean()) < 40                              # ~31 in-radius neighbours

    A = torch.rand(N, 16, device=DEV, generator=None)
    B_ = torch.rand(N, 16, device=DEV)
    G = torch.randn(N, 16, 3, device=DEV)
    Ar = A.clone().requires_grad_(True)
    GA = sphops.gradient(x, v, Ar, h, grid)
    GA.backward(G)
    lhs = (G.double() * GA.detach().double()).sum()
    rhs = (Ar.grad.double() * A.double()).sum()
    # both sides are heavily cancelling sums of ~5e7 terms: the bound scales with sum|terms|, not |sum|
    scale = (G.double().abs() * GA.detach().double().abs()).sum()
    assert abs(float(lhs - rhs)) <= 1e-6 * float(scale), (float(lhs), float(rhs), float(scale))
    GB = sphops.gradient(x, v, B_, h, grid)
    GAB = sphops.gradient(x, v, 2.0 * A + B_, h, grid)
    assert rel_err(GAB, 2.0 * GA.detach() + GB) <= 2e-5
    const = torch.full((N, 4), 3.25, device=DEV)
    assert float(sphops.gradient(x, v, const, h, grid).abs().max()) == 0.0
    ones = torch.ones(N, 1, device=DEV)
    S = sphops.blur(x, v, ones, h, grid)
    assert 0.8 < float(S.mean()) < 1.2
