"""Import alias: `model.*` of the reference resolves to latent_feature_grid_compression_b200.model.* (see INTEGRATION.md)."""
