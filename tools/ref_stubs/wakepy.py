"""Empty stand-in for `wakepy` (imported at the top of the reference's upscale_processing.py, unused on the hot path)."""
keep = None
