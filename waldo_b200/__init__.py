"""waldo_b200 -- B200-native (sm_100a) warp + composite hot path of WALDO behind the reference's module interfaces.

See DESIGN.md for the path and its boundary, include/waldo_b200.h for the C ABI, INTEGRATION.md for how the
reference binds to it.
"""
from .modules import (TPSWarp, InverseWarp, Warper, compute_occ, decode_output, estimate_alpha_grid_occ, alpha_masks,
                      wif_fuse, pack_input, frames_to_u8, blur, layer_entropy, pose_distance_losses, obj_flow_loss, wif_to_emb, conv3x3, Conv3x3, get_grid, get_gaussian_kernel, kernel_distance)
from . import functional
from .functional import set_deterministic, is_deterministic
from .feed import DevicePrefetcher
from .graphs import GraphedDecode

__all__ = ["TPSWarp", "InverseWarp", "Warper", "compute_occ", "decode_output", "estimate_alpha_grid_occ", "alpha_masks",
           "wif_fuse", "pack_input", "frames_to_u8", "blur", "layer_entropy", "pose_distance_losses", "obj_flow_loss", "wif_to_emb", "conv3x3", "Conv3x3", "get_grid", "get_gaussian_kernel", "kernel_distance", "functional", "DevicePrefetcher", "GraphedDecode",
           "set_deterministic", "is_deterministic"]
