"""Golden vectors for the rollout drivers of SURVEY.md 8f: nca.sample_plane_wrapped and nca.sample_mesh
(+ diffuse), produced by the UNMODIFIED reference under Numba's CUDA simulator (see make_golden.py)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

sphops, nca = ref_shim.load_reference()
import numpy as np  # noqa: E402
import torch  # noqa: E402
from make_golden import jittered_lattice, load_weights, np32, save  # noqa: E402


def model_with(weights, **kw):
    m = nca.SPHNCA(16, 0.5, **kw)
    W1, b1, W2, b2 = load_weights(weights)
    with torch.no_grad():
        m.model[0].weight.copy_(W1); m.model[0].bias.copy_(b1)
        m.model[2].weight.copy_(W2); m.model[2].bias.copy_(b2)
    return m, (W1, b1, W2, b2)


def wrapped_case():
    gen = torch.Generator().manual_seed(21)
    nca.USE_3D = True
    x2 = jittered_lattice(12, 2, gen, jitter=0.6)
    N = x2.shape[0]
    sections = (N // 2, N - N // 2)
    h, dims = 0.4, 7                                  # wrapped points reach out to +-1.4
    m, W = model_with("zebra", update_rule="gated", use_alpha=False, normalize_perception=2.5)
    A0 = torch.rand(N, 16, generator=gen)
    gmin, gsize = torch.tensor([-1.0, -1.0]), torch.tensor([2.0, 2.0])
    torch.manual_seed(5)
    with torch.no_grad():
        loss, x_out, out, x_w, out_w = nca.sample_plane_wrapped(m, x2, A0, sections, h, dims, gmin, gsize, out_steps=True, iter_n=3)
    torch.manual_seed(5)
    masks = [np32(torch.rand(x_w.shape[0]) <= 0.5) for _ in range(3)]
    save("wrapped_zebra_k3", x=np32(x2), sections=np.asarray(sections, np.int64), h=np.float64(h), dims=np.int64(dims),
         W1=np32(W[0]), b1=np32(W[1]), W2=np32(W[2]), b2=np32(W[3]), A0=np32(A0), normalize_perception=np.float64(2.5),
         x_wrapped=np32(x_w), states=np.stack([np32(s) for s in out]), states_wrapped_last=np32(out_w[-1]),
         masks=np.stack(masks))


def wrapped_periodic_case():
    """The periodic plane as train.py / test.py set it up (train.py:69 DIMS = ceil(2 / H), so dims * h == the plane size and the
    hash grid's modulo wrap IS the plane's period): forward states AND gradients through the ghost refresh."""
    gen = torch.Generator().manual_seed(23)
    nca.USE_3D = True
    x2 = jittered_lattice(12, 2, gen, jitter=0.6)
    N = x2.shape[0]
    sections = (N // 2, N - N // 2)
    h = 0.4
    dims = int(np.ceil(2 / h))                        # 5
    m, W = model_with("zebra", update_rule="gated", use_alpha=False, normalize_perception=2.5)
    A0 = torch.rand(N, 16, generator=gen).requires_grad_(True)
    gmin, gsize = torch.tensor([-1.0, -1.0]), torch.tensor([2.0, 2.0])
    torch.manual_seed(5)
    loss, x_out, out, x_w, out_w = nca.sample_plane_wrapped(m, x2, A0, sections, h, dims, gmin, gsize, out_steps=True, iter_n=3)
    L = (out[-1][:, :4] ** 2).mean() + 0.1 * (out[1][:, :4] ** 2).mean()
    L.backward()
    torch.manual_seed(5)
    masks = [np32(torch.rand(x_w.shape[0]) <= 0.5) for _ in range(3)]
    P = dict(m.named_parameters())
    save("wrapped_periodic_k3", x=np32(x2), sections=np.asarray(sections, np.int64), h=np.float64(h), dims=np.int64(dims),
         W1=np32(W[0]), b1=np32(W[1]), W2=np32(W[2]), b2=np32(W[3]), A0=np32(A0.detach()), normalize_perception=np.float64(2.5),
         x_wrapped=np32(x_w), states=np.stack([np32(s.detach()) for s in out]), states_wrapped_last=np32(out_w[-1].detach()),
         masks=np.stack(masks), loss=np.float64(L.item()), dA0=np32(A0.grad), dW1=np32(P["model.0.weight"].grad),
         db1=np32(P["model.0.bias"].grad), dW2=np32(P["model.2.weight"].grad), db2=np32(P["model.2.bias"].grad))


def mesh_case():
    gen = torch.Generator().manual_seed(22)
    nca.USE_3D = True
    n = 260
    # points on the unit sphere, scaled into [-0.9, 0.9]^3
    p = torch.randn(n, 3, generator=gen)
    x = 0.9 * p / p.norm(dim=-1, keepdim=True)
    Nrm = x / x.norm(dim=-1, keepdim=True)
    T = nca.orthogonalize(Nrm, torch.randn(n, 3, generator=gen))
    h, dims = 0.35, 6
    m, W = model_with("gecko", update_rule="gated", use_alpha=True, normalize_perception=1.0 / h)
    A0 = 0.5 + 0.5 * torch.rand(n, 16, generator=gen)
    T0 = T.clone()
    torch.manual_seed(6)
    with torch.no_grad():
        loss, x_out, out = nca.sample_mesh(m, x, A0, Nrm, T, (n,), h, dims, out_steps=True, iter_n=2)
    torch.manual_seed(6)
    masks = [np32(torch.rand(n) <= 0.5) for _ in range(2)]
    save("mesh_gecko_k2", x=np32(x), N=np32(Nrm), T0=np32(T0), T_final=np32(T), h=np.float64(h), dims=np.int64(dims),
         W1=np32(W[0]), b1=np32(W[1]), W2=np32(W[2]), b2=np32(W[3]), A0=np32(A0), normalize_perception=np.float64(1.0 / h),
         states=np.stack([np32(s) for s in out]), masks=np.stack(masks))


if __name__ == "__main__":
    torch.set_num_threads(1)
    only = sys.argv[1:]
    for name, fn in (("wrapped", wrapped_case), ("wrapped_periodic", wrapped_periodic_case), ("mesh", mesh_case)):
        if not only or name in only:
            fn()
