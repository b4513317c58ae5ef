"""Drop-in counterpart of the reference's ``data`` package."""
