"""Drop-in counterpart of the reference's ``visualization`` package (reconstruction + config reader only)."""
