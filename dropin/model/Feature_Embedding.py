"""Alias of latent_feature_grid_compression_b200.model.Feature_Embedding under the reference's module name."""
from latent_feature_grid_compression_b200.model.Feature_Embedding import *  # noqa: F401,F403
from latent_feature_grid_compression_b200.model import Feature_Embedding as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
