"""msfwsi_b200 -- B200-native (sm_100a) SSL head + loss hot path of MSF-WSI.

The package name drops the hyphen of the repository name ("msf-wsi") because a hyphen is
not a legal Python identifier.  Public surface:

  MSFWSI, ssl_loss                      drop-in for src/models/backbone.py::MSFWSI + the loss block of train()
  resnet18 / resnet34                   encoders with return_features=True (PyTorch/cuDNN; not a CUDA target)
  ops.gather_concat / cosine_loss / infonce_loss / crop_resample / bn_act2d / EmaUpdater   the C-ABI operators
  FusedAdam                             torch.optim.Adam with the step (+ unscale, overflow skip, EMA) in one launch
  checkpoint                            the reference driver's checkpoint layout (save / resume / fine-tune key surgery)
  GraphedStep                           one whole training step captured in ONE CUDA graph and replayed (single GPU)

Importing the package never touches the GPU; the first operator call loads
lib/libmsfwsi_b200.so and raises if it has not been built (no CPU fallback).
"""
from . import ops  # noqa: F401
from .module import DEFAULT_FUSER_WEIGHTS, MSFWSI, TCLinear, bind_optimizer, make_predictor, make_projector, ssl_loss  # noqa: F401
from . import checkpoint  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .graph import GraphedStep  # noqa: F401
from .resnet import resnet18, resnet34  # noqa: F401

__all__ = ["MSFWSI", "ssl_loss", "resnet18", "resnet34", "ops", "make_projector", "make_predictor",
           "DEFAULT_FUSER_WEIGHTS", "FusedAdam", "bind_optimizer", "checkpoint", "GraphedStep"]
