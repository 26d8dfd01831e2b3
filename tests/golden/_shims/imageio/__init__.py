"""Import shim for the fixture generator: the render path never calls imageio; the dataset loaders
call `imageio.imread(path[, pilmode="RGB"])`, which is restated with PIL (same decoded pixels for PNG)."""
import numpy as np
from PIL import Image


def imread(uri, pilmode=None, ignoregamma=None, **kwargs):
    img = Image.open(uri)
    if pilmode is not None:
        img = img.convert(pilmode)
    return np.array(img)
