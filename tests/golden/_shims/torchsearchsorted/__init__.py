"""Import shim used ONLY by make_golden.py: the reference imports the un-vendored
`torchsearchsorted` extension (nerf-pytorch/requirements.txt:9).  Its documented contract is
numpy.searchsorted semantics per row, which torch.searchsorted implements."""
import torch


def searchsorted(a, v, side="left"):
    return torch.searchsorted(a, v, right=(side == "right"))
