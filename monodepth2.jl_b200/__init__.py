"""B200-native view-synthesis loss path of pxl-th/Monodepth2.jl: host-side mirror of the
reference's loss/geometry operator interface over the C-ABI library csrc/libmd2_b200.so
(hand-written sm_100a CUDA kernels, forward + backward).  CUDA only, no CPU fallback."""
from ._lib import Context, Md2Error, load_library, EXPORTS, LIB_PATH  # noqa: F401
from .ops import (SSIM, Backproject, Project, _apply_mask, automasking_loss, composeT,  # noqa: F401
                  disparity_to_depth, grid_sample, hat, photometric_loss, prediction_loss, smooth_loss,
                  so3_exp_map, upsample_bilinear)
from .training import (Adam, AsyncViz, HostViewSynthesisLoss, Params, Pose, TrainCache, simple_depth_loss, slow_depth,  # noqa: F401
                       train_loss, view_synthesis_loss, warp)

from .train_step import DataParallelTrainer, GradientBuckets, StandInModel, make_training_setup  # noqa: F401

__version__ = "0.1.0"
