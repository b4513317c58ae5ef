"""Alias of latent_feature_grid_compression_b200.wavelet_transform.Torch_Wavelet_Transform under the reference's module name."""
from latent_feature_grid_compression_b200.wavelet_transform.Torch_Wavelet_Transform import *  # noqa: F401,F403
from latent_feature_grid_compression_b200.wavelet_transform import Torch_Wavelet_Transform as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
