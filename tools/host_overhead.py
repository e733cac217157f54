"""Host (Python + ctypes) enqueue time vs GPU time of one bench step (development aid): if the enqueue time of a
step is not well below its GPU time, the step is launch-bound on a slow host."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sph-nca_b200"))
import torch
import nca, multigpu
from bench import make_inputs, WORKLOADS, load_weights, bench_loss
dev = torch.device("cuda", 0)
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "train2d_b8x65k_k64"]
nca.USE_3D = True
x, A, sections, h, dims = make_inputs(wl, 0)
K = wl["k_nca"]
model = multigpu.DataParallelSPHNCA(16, 0.5, update_rule="gated", normalize_grad=True, use_alpha=False, normalize_perception=1.0 / h).to(dev)
with torch.no_grad():
    for t, src in zip((model.model[0].weight, model.model[0].bias, model.model[2].weight, model.model[2].bias), load_weights()):
        t.copy_(src)
opt = torch.optim.Adam(model.parameters(), lr=1e-5)
xd, Ad = x.to(dev), A.to(dev)
def step():
    return nca.sample_plane(model, xd, Ad, sections, h, dims, iter_n=K, loss_fn=bench_loss, img=None, optimizer=opt)
for _ in range(3): step()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"host enqueue {1e3 * (t1 - t0):.1f} ms (incl. the loss .item() sync), total {1e3 * (t2 - t0):.1f} ms")
# forward only, no sync inside
xp = torch.nn.functional.pad(xd, (0, 1))
grid, v = model.init_grid(xp, sections, h, dims)
A0 = Ad.clone().requires_grad_(True)
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); out = model.rollout(xp, v, A0, h, grid, K); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    loss = (out[-1][:, :4] ** 2).mean()
    torch.cuda.synchronize()
    t3 = time.perf_counter(); loss.backward(); t4 = time.perf_counter(); torch.cuda.synchronize(); t5 = time.perf_counter()
    print(f"rollout fwd: host {1e3 * (t1 - t0):.1f} ms, gpu-complete {1e3 * (t2 - t0):.1f} ms | bwd: host {1e3 * (t4 - t3):.1f} ms, gpu-complete {1e3 * (t5 - t3):.1f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); out = model.rollout(xp, v, A0, h, grid, K); (out[-1][:, :4] ** 2).mean().backward(); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
