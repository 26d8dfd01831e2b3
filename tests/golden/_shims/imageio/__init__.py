"""Empty import shim (the render path never calls imageio)."""
