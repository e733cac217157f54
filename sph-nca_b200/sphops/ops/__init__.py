from . import func

volume = func.Volume.apply
gradient = func.Gradient.apply
divergence = func.Divergence.apply
count = func.Count.apply
blur = func.Blur.apply
