"""Import alias: `wavelet_transform.*` of the reference resolves to latent_feature_grid_compression_b200.wavelet_transform.* (see INTEGRATION.md)."""
