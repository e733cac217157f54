"""Stress run of the pipelined tcgen05 kernels and the fire-list path (development aid): many launches with random row counts
and fire rates through the C ABI; every list-mode result must equal the all-rows result on the listed rows bit for bit
(forward) / to summation order (weight gradient), and repeated launches must be bit-identical."""
import ctypes as C, sys, os, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sph-nca_b200"))
import torch
from sphops import _lib, fused
from bench import load_weights
dev = "cuda"
lib, ptr, check = _lib.lib, _lib.ptr, _lib.check
W = [t.to(dev) for t in load_weights()]
p = fused._step_params(*W, "gated", False, 1.0, 1.0)
st = _lib.stream_ptr()
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
f32 = dict(dtype=torch.float32, device=dev)
bad = 0
for it in range(iters):
    N = rng.choice([1, 7, 127, 128, 129, 1000, 4096, 18944, 18945, 50000, 150000, 300001])
    rate = rng.choice([0.0, 0.03, 0.5, 0.5, 0.97, 1.0])
    g = torch.Generator(device=dev).manual_seed(it)
    A = torch.rand(N, 16, generator=g, **f32); Y = torch.randn(N, 32, generator=g, **f32)
    fire = (torch.rand(N, generator=g, device=dev) < rate).to(torch.uint8)
    flags0 = torch.ones(N, dtype=torch.uint8, device=dev)
    # ---- all rows
    An = torch.empty(N, 16, **f32); O = torch.empty(N, 36, **f32); al = torch.empty(N, dtype=torch.uint8, device=dev)
    rl = torch.empty(N, 8, dtype=torch.int32, device=dev); fl = flags0.clone()
    check(lib.sphnca_step_mlp(N, ptr(A), ptr(Y), ptr(fire), C.byref(p), ptr(An), ptr(al), ptr(O), ptr(fl), ptr(rl), st))
    # ---- fire list
    rows = torch.empty(N, dtype=torch.int32, device=dev); cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    An2 = torch.full((N, 16), -3.0, **f32); al2 = torch.empty(N, dtype=torch.uint8, device=dev)
    nb = C.c_size_t(0); check(lib.sphnca_fire_list_workspace_bytes(N, C.byref(nb)))
    ws = torch.zeros(nb.value, dtype=torch.uint8, device=dev)
    check(lib.sphnca_fire_list(N, ptr(fire), ptr(A), 0, ptr(rows), ptr(cnt), ptr(An2), ptr(al2), None, None, 0, ptr(ws), nb.value, st))
    O2 = torch.zeros(N, 36, **f32); rl2 = torch.zeros(N, 8, dtype=torch.int32, device=dev); fl2 = flags0.clone()
    Ac = torch.zeros(N, 16, **f32); Yc = torch.zeros(N, 32, **f32)
    check(lib.sphnca_step_mlp_list(N, ptr(rows), ptr(cnt), ptr(A), ptr(Y), C.byref(p), ptr(An2), ptr(al2), ptr(O2), ptr(fl2), ptr(rl2),
                                   ptr(Ac), ptr(Yc), 0, st))
    n = int(cnt.item()); r = rows[:n].long()
    why = []
    def chk(cond, what):
        if not cond: why.append(what)
        return cond
    ok = chk(n == int(fire.sum()), "count") & chk(torch.equal(An2, An), "A_new") & chk(torch.equal(O2[:n], O[r]), "O") \
        & chk(torch.equal(rl2[:n], rl[r]), "relu") & chk(torch.equal(Ac[:n], A[r]), "A_c") & chk(torch.equal(Yc[:n], Y[r]), "Y_c") \
        & chk(torch.equal(fl2, fl), "flags")
    # ---- backward: all rows vs list
    fl |= 2; fl2 |= 2
    dAn = torch.randn(N, 16, generator=g, **f32)
    dO = torch.empty(N, 36, **f32); dY = torch.empty(N, 32, **f32); dAd = torch.empty(N, 16, **f32)
    check(lib.sphnca_step_bwd(N, ptr(A), ptr(Y), ptr(O), ptr(fl), ptr(rl), ptr(dAn), C.byref(p), ptr(dO), ptr(dY), ptr(dAd), None, None, st))
    dO2 = torch.zeros(N, 36, **f32); dY2 = torch.zeros(N, 32, **f32); dAd2 = torch.zeros(N, 16, **f32)
    check(lib.sphnca_step_bwd_unfired(N, ptr(fl2), ptr(dAn), ptr(dY2), ptr(dAd2), st))
    check(lib.sphnca_step_bwd_list(N, ptr(rows), ptr(cnt), ptr(Ac), ptr(O2), ptr(fl2), ptr(rl2), ptr(dAn), C.byref(p), ptr(dO2), ptr(dY2), ptr(dAd2), st))
    ok = ok & chk(torch.equal(dY2, dY), "dY") & chk(torch.equal(dAd2, dAd), "dAdir") & chk(torch.equal(dO2[:n], dO[r]), "dO")
    # ---- weight gradient: list vs all rows (not-fired rows have dO == 0), and twice the same bits
    nbw = C.c_size_t(0); check(lib.sphnca_mlp_wgrad_workspace_bytes(N, C.byref(nbw)))
    wsw = torch.empty(nbw.value, dtype=torch.uint8, device=dev)
    outs = []
    for mode in ("all", "list", "list"):
        dW = [torch.zeros_like(t) for t in W]
        if mode == "all":
            check(lib.sphnca_mlp_wgrad(N, ptr(A), ptr(Y), ptr(dO), None, None, C.byref(p), *[ptr(t) for t in dW], ptr(wsw), nbw.value, st))
        else:
            check(lib.sphnca_mlp_wgrad_list(N, ptr(cnt), ptr(Ac), ptr(Yc), ptr(dO2), C.byref(p), *[ptr(t) for t in dW], ptr(wsw), nbw.value, st))
        outs.append(dW)
    for a, b, c_ in zip(*outs):
        scale = float(a.abs().max()) or 1.0
        # random-sign test data: the sums over ~150 k rows cancel to ~sqrt(n) of their terms, so two summation orders differ by
        # a few 1e-5 of the largest entry (measured: 2.5e-5 at N = 300,001); the bitwise checks are the ones that catch races
        ok = ok & chk(float((a - b).abs().max()) <= 1e-4 * scale, f"wgrad all-vs-list {float((a - b).abs().max()) / scale:.2e}") \
            & chk(torch.equal(b, c_), "wgrad list twice")
    torch.cuda.synchronize()
    if not ok:
        bad += 1
        print(f"MISMATCH it={it} N={N} rate={rate} fired={n}: {why}")
print(f"stress: {iters} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
