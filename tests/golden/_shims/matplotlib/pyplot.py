"""Empty import shim."""
