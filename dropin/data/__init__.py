"""Import alias: `data.*` of the reference resolves to latent_feature_grid_compression_b200.data.* (see INTEGRATION.md)."""
