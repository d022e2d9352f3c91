"""u96_slam_b200 -- B200-native dense-stereo front end of U96-SLAM (rect -> x-Sobel -> SAD BM ->
disparity -> 3-D) behind the C ABI in include/u96_stereo.h.  CUDA only: importing the bindings
fails loudly when libu96stereo.so is missing; there is no CPU fallback."""
from .stereo import (Fpga, StereoBM, StereoFrontEnd, U96Error, PROFILE_OPENCV, PROFILE_RTL,  # noqa: F401
                     SHIPPED_RECT_PARAMS, LOCAL_TRANSFORM, UVC_BM, UVC_RECT, UVC_XSBL, lib_path, load_library)
from .synth import synth_pair, synth_batch, identity_rect_params  # noqa: F401
from .formats import (read_dat, write_dat, rect_params_from_calibration, load_projection_kitti,  # noqa: F401
                      load_projection_opencv_yml)
