"""CPU-side checks: the C-ABI library loads and exports every symbol include/sphnca.h declares,
the host-side mirror of the reference interface has the reference's names, and the product
refuses to run without CUDA (no fallback).  No compute calls are made here."""
import os
import re

import pytest
import torch

from tests.util import ROOT


def test_header_symbols_are_exported():
    import sphops._lib as L
    header = open(os.path.join(ROOT, "include", "sphnca.h")).read()
    declared = set(re.findall(r"\b(sphnca_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert set(L.DECLARED_SYMBOLS) == declared, (set(L.DECLARED_SYMBOLS) ^ declared)
    assert L.MISSING_SYMBOLS == (), f"not exported by libsphnca.so: {L.MISSING_SYMBOLS}"
    assert L.lib.sphnca_abi_version() == 3


def test_public_names_match_reference():
    import sphops
    for name in ("initialize_hashgrid", "volume", "gradient", "divergence", "count", "blur"):
        assert callable(getattr(sphops, name))
    from sphops.ops import func
    for cls in ("Volume", "Gradient", "Divergence", "Count", "Blur"):
        assert issubclass(getattr(func, cls), torch.autograd.Function)
    from sphops.utils import pack, unpack
    a, b = torch.arange(6.).view(3, 2), torch.arange(4.).view(2, 2)
    packed, sections = pack(a, b)
    assert sections == [3, 2] and packed.shape == (5, 2)
    ua, ub = unpack(packed, sections)
    assert torch.equal(ua, a) and torch.equal(ub, b)
    from sphops.preprocess import Hashgrid
    from sphops.hashgrid import single_impl
    assert single_impl.Hashgrid is Hashgrid


def test_nca_api_and_checkpoint_keys():
    import nca
    m = nca.SPHNCA(16, 0.5)
    assert sorted(m.state_dict().keys()) == ["model.0.bias", "model.0.weight", "model.2.bias", "model.2.weight"]
    assert m.model[0].weight.shape == (256, 48) and m.model[2].weight.shape == (33, 256)
    assert sum(p.numel() for p in m.parameters()) == 21025
    o = nca.SPHNCA(16, 0.5, update_rule="orig")
    assert o.model[2].weight.shape == (16, 256) and float(o.model[2].weight.abs().sum()) == 0.0
    for name in ("sample_plane", "cell_activity", "default_feature_processs", "USE_3D"):
        assert hasattr(nca, name)
    feats = [(torch.ones(5, 16), torch.arange(5 * 16 * 3.).view(5, 16, 3))]
    y = nca.default_feature_processs(feats)
    assert y.shape == (5, 48) and torch.equal(y[:, 16:32], feats[0][1][..., 0]) and torch.equal(y[:, 32:], feats[0][1][..., 1])


def test_no_cpu_fallback():
    import sphops
    x = torch.rand(10, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        sphops.initialize_hashgrid(x, (10,), 0.5, 4)


def test_bad_arguments_are_rejected_by_the_c_abi():
    import ctypes as C
    import sphops._lib as L
    n = C.c_size_t(0)
    assert L.lib.sphnca_hashgrid_workspace_bytes(10, 1, 0, C.byref(n)) == -1
    assert b"bad sizes" in L.lib.sphnca_last_error()
    assert L.lib.sphnca_hashgrid_workspace_bytes(1000, 2, 125, C.byref(n)) == 0 and n.value > 0
    # D outside {2,3} is refused before any CUDA call is made
    sec = (C.c_int64 * 1)(4); dims = (C.c_int64 * 3)(2, 2, 2)
    rc = L.lib.sphnca_hashgrid_build(None, 4, 4, sec, 1, 0.5, dims, None, None, None, None, None, None, None, None, 0, None)
    assert rc == -1 and b"D must be 2 or 3" in L.lib.sphnca_last_error()


def test_workspace_sizes_and_struct_layout_without_a_gpu():
    """Host-only entries: the fire-list workspace holds one 8-byte state word per block of 1024 rows, one per group of 256
    blocks and the two counters (include/sphnca.h: zero before the first use); struct sphnca_step_params is 56 bytes with
    dy_interleaved behind orig_rate (ABI 3)."""
    import ctypes as C
    import sphops._lib as L
    for n in (0, 1, 1024, 1025, 524_288, 262_145 * 1024, 2**31 - 1):
        nb = C.c_size_t(0)
        L.check(L.lib.sphnca_fire_list_workspace_bytes(n, C.byref(nb)))
        blocks = (n + 1023) // 1024
        assert nb.value == (blocks + (blocks + 255) // 256 + 1) * 8 and nb.value % 8 == 0, (n, nb.value)
    assert L.lib.sphnca_fire_list_workspace_bytes(-1, C.byref(nb)) != 0
    assert C.sizeof(L.CStepParams) == 56
    assert L.CStepParams.dy_interleaved.offset == 48 and L.CStepParams.orig_rate.offset == 44
    p = L.CStepParams()
    assert p.dy_interleaved == 0            # every entry but the fire-list backward pair ignores it; default = plain layout
