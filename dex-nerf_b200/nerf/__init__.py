"""`nerf` - B200-native drop-in for the Python namespace of Dex-NeRF's nerf-pytorch core
(reference: nerf-pytorch/nerf/__init__.py:1-8 flattens the same sub-modules).

Everything numeric runs in hand-written sm_100a CUDA kernels behind the C ABI of
include/dexnerf.h (loaded with ctypes from dex-nerf_b200/lib/libdexnerf.so).  There is no CPU
path: CPU tensors raise ValueError, a missing library raises DexNerfError."""
from . import models
from ._lib import DexNerfError
from .cfgnode import CfgNode
from .models import *  # noqa: F401,F403
from .models import (FlexibleNeRFModel, MultiHeadNeRFModel, PaperNeRFModel, ReplicateNeRFModel,
                     VeryTinyNeRFModel)
from .nerf_helpers import (cumprod_exclusive, gather_cdf_util, get_embedding_function, get_minibatches,
                           get_ray_bundle, img2mse, meshgrid_xy, mse2psnr, ndc_rays, positional_encoding,
                           sample_pdf_2)
from .train_utils import (get_precision, predict_and_render_radiance, render_camera, run_network,
                          run_one_iter_of_nerf, sample_pdf, set_precision)
from .sharding import SharedFrame, allreduce_gradients, gather_rows, row_block
from .training import Trainer, learning_rate, train_step
from .datasets import load_blender_data, load_llff_data, load_messytable_data
from .eval_utils import (cast_to_image, compute_err_metric, depth_error_img, dex_depth_error_metrics, pose_spherical, render_path,
                         render_poses_spherical, select_dex_threshold, world2cam_from_blender_pose)
from .volume_rendering_utils import volume_render_radiance_field
from . import cache_utils
