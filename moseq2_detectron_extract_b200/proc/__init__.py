from .proc import (clean_frames, crop_and_rotate_frame, crop_and_rotate_frames_batch, filter_angles,  # noqa: F401
                   find_invalid_pixels, flips_from_keypoints, get_frame_features, im_moment_features,
                   instances_to_features, iterative_filter_angles, mask_and_keypoints_from_model_output,
                   prep_raw_frames, scale_raw_frames, clamp_angles_deg, InvalidPixelsError)
from .roi import apply_roi, get_bbox, get_bground_im, get_roi, plane_fit3, plane_ransac  # noqa: F401
from .scalars import compute_scalars, scalar_attributes  # noqa: F401
from .keypoints import keypoints_to_dict, keypoint_attributes, rotate_points, rotate_points_batch  # noqa: F401
from .util import convert_pxs_to_mm, select_strel  # noqa: F401
