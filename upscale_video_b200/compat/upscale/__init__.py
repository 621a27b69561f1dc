"""Stand-in for the reference's ``upscale`` package: only the hot-path module is provided (see ../README.md)."""
