"""Drop-in counterpart of the reference's ``wavelet_transform`` package."""
