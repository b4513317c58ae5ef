"""Import alias: `visualization.*` of the reference resolves to latent_feature_grid_compression_b200.visualization.* (see INTEGRATION.md)."""
