""""train.py and test.py run unchanged" as evidence (BASELINE.json north_star): the reference's OWN driver
scripts and `losses.py` -- the unmodified files staged under git-ignored `oracle/_ref/code` by `oracle/stage_reference.py` --
are executed on top of THIS repo's `sphops` + `nca` (`sph-nca_b200/` ahead of the staged reference on sys.path, so
`import nca` / `import sphops` resolve here and `commons`, `losses` resolve to the reference).

* CPU (no GPU needed): the reference's `losses.get_mse_loss` + this repo's `nca.Pool` / `SPHNCA` -- import resolution, the
  loss on pool samples, `replace_worst`, `update`;
* GPU: `train.py` for 3 iterations (the gecko example's flags with a local image instead of the emoji download, and the
  texture example's `--use_alpha false --wrap true`), then `test.py` on the checkpoint format train.py writes.
Skipped when oracle/_ref is not staged.  Only tests may execute anything under oracle/.
"""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sph-nca_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "code")
staged = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "sphops")), reason="oracle/_ref/code is not staged")

RUNNER = """
import os, runpy, sys
pkg, ref, script = sys.argv[1:4]
sys.argv = [script] + sys.argv[4:]
sys.path[:0] = [pkg, ref]
os.chdir(ref)
import nca, sphops
assert os.path.realpath(nca.__file__).startswith(os.path.realpath(pkg)), nca.__file__
assert os.path.realpath(sphops.__file__).startswith(os.path.realpath(pkg)), sphops.__file__
runpy.run_path(os.path.join(ref, script), run_name="__main__")
print("__RAN__", script, "with", nca.__file__)
"""


def run_reference_script(script, argv, timeout=600):
    r = subprocess.run([sys.executable, "-c", RUNNER, PKG, REF, script] + argv, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, f"{script} failed:\n{r.stdout[-2000:]}\n{r.stderr[-3000:]}"
    assert "__RAN__" in r.stdout
    return r.stdout


@staged
def test_reference_losses_and_pool_on_this_nca_cpu():
    code = """
import os, sys
pkg, ref = sys.argv[1:3]
sys.path[:0] = [pkg, ref]
import torch
import nca, losses
from commons import geometry
assert os.path.realpath(nca.__file__).startswith(os.path.realpath(pkg))
assert os.path.realpath(losses.__file__).startswith(os.path.realpath(ref))
M = 16
GSHAPE, GMIN, GSIZE = torch.tensor([M, M]), torch.tensor([-1, -1]), torch.tensor([2, 2])
x = geometry.grange(GSHAPE, GMIN, GSIZE).view(-1, 2)
A = torch.zeros(x.shape[0], 16)
nca.add_radial_seed(x, A, center=GMIN + GSIZE / 2, R=0.3)
model = nca.SPHNCA(16, 0.5, use_alpha=True, normalize_grad=True, normalize_perception=1 / 0.2)
img = torch.rand(12, 12, 4)
cfg = dict(IMAGE_SCALE=12 / M, GMIN=GMIN, GSIZE=GSIZE, LOSS_WEIGHT_OVERFLOW=0.05)
loss_fn = losses.get_mse_loss(model, img, cfg)
pool = nca.Pool((x, A), 8)
x0, A0, sections, idx = pool.sample(4, "cpu", replace_worst=True, loss_fn=loss_fn, img=img)
assert x0.shape == (4 * M * M, 2) and A0.shape == (4 * M * M, 16) and sections == (M * M,) * 4
l = loss_fn(x0[:M * M], A0[:M * M], img)
assert torch.isfinite(l) and l.dim() == 0
pool.update(x0, A0 + 1, idx)
assert torch.equal(pool.A[idx[1]], (A0 + 1).view(4, -1, 16)[1])
print("OK")
"""
    r = subprocess.run([sys.executable, "-c", code, PKG, REF], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


@staged
@pytest.mark.gpu
@pytest.mark.parametrize("flags", [
    # the shipped gecko example (train-example.sh:8-13) with a local image instead of the emoji download, small sizes
    ["--loss", "mse_simple", "--loss_weight_overflow", "0.05", "--initial_feature", "radial", "--initial_feature_radius", "0.16",
     "--use_alpha", "true", "--alpha_premultiply", "true", "--wrap", "false", "--h", "0.1", "--image_size", "48", "--target_size", "40",
     "--steps_range", "4,6", "--steps_increment", "1", "--nca_update", "gated"],
    # the texture examples' model flags (train-example.sh:18,25: no alpha channel, periodic plane), mse loss
    ["--loss", "mse_simple", "--initial_feature", "random", "--use_alpha", "false", "--wrap", "true", "--h", "0.1",
     "--image_size", "40", "--target_size", "40", "--steps_range", "4,6", "--steps_increment", "1", "--nca_update", "gated"],
], ids=["gecko_flags_alpha", "texture_flags_noalpha_wrapped"])
def test_reference_train_py_runs_unchanged(flags, tmp_path):
    out = run_reference_script("train.py", ["--gpu_index", "0", "--seed", "1", "--img", os.path.join(REF, "data", "striped_0005.jpg"),
                                            "--lr", "3e-3", "--training_iter", "3", "--batch_size", "4", "--pool_size", "8",
                                            "--output_dir", str(tmp_path)] + flags)
    assert "Done!" in out and "#params = 21025" in out


@staged
@pytest.mark.gpu
@pytest.mark.parametrize("wrap", ["true", "false"])
def test_reference_test_py_runs_unchanged(wrap, tmp_path):
    sys.path[:0] = [PKG] if PKG not in sys.path else []
    import nca
    m = nca.SPHNCA(16, 0.5, use_alpha=False, normalize_perception=1 / 0.1)
    ckpt = tmp_path / "ckpt.pt"
    torch.save({"model": m.state_dict(), "epoch": 0}, ckpt)              # the format train.py:311-317 writes
    out = run_reference_script("test.py", ["--gpu_index", "0", "--seed", "1", "--checkpoint", str(ckpt), "--initial_feature", "random",
                                           "--use_alpha", "false", "--wrap", wrap, "--h", "0.1", "--image_size", "32", "--steps", "6",
                                           "--nca_update", "gated", "--output_dir", str(tmp_path / "out")])
    files = [f for _, _, fs in os.walk(tmp_path / "out") for f in fs]
    assert sum(f.endswith(".png") for f in files) == 7 and any(f.endswith(".pt") for f in files), (files, out[-500:])
