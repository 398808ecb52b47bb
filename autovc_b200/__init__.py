"""autovc_b200 — B200-native (sm_100a) drop-in for the AutoVC Generator training hot path.

Public surface mirrors the reference's modules for this path:
  autovc_b200.model_vc_mel.Generator     <- model_vc_mel.Generator
  autovc_b200.model_vc_stft.GeneratorSTFT <- model_vc_stft.GeneratorSTFT
  autovc_b200.model_vc_wav.GeneratorWav   <- model_vc_wav.GeneratorWav (Conv-TasNet filterbanks around the AutoVC core)
  autovc_b200.make_spect.Spect / logmel   <- make_spect.Spect (spmel branch)
  autovc_b200.data_loader.get_loader      <- data_loader.get_loader (corpus resident in HBM, one launch per batch)
  autovc_b200.model_bl.D_VECTOR           <- model_bl.D_VECTOR (speaker encoder, inference)
  autovc_b200.optim.FusedAdam             <- torch.optim.Adam as solver_encoder.py:130 configures it (one-launch step)
  autovc_b200.solver                      <- the step maths of solver_encoder.Solver.train + data parallelism
"""
from ._lib import AvcError, LIB_PATH, launch_count, load  # noqa: F401
from .model_vc_mel import Generator  # noqa: F401
from .model_vc_stft import GeneratorSTFT  # noqa: F401
from .model_vc_wav import GeneratorWav  # noqa: F401
from .optim import FusedAdam  # noqa: F401

__all__ = ["Generator", "GeneratorSTFT", "GeneratorWav", "FusedAdam", "AvcError", "load", "launch_count", "LIB_PATH"]
