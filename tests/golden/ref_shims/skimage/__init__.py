"""Import-only stub (skimage is absent offline; never called on the hot path)."""
