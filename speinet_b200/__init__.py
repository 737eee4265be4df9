"""speinet_b200: B200-native (sm_100a) SearchTransfer hot path of yangt1013/SPEINet.

Only what the path needs: the CUDA kernels + C-ABI (`csrc/`, `include/speinet_b200.h`), and the
host-side mirror of the reference interface (`SearchTransfer`, `SelfTransfer`, `fuse_level`,
`install`).  No CPU path, no PyTorch fallback.
"""
from ._lib import LIB_PATH, load as load_library  # noqa: F401
from .search_transfer import SearchTransfer, SelfTransfer, search_transfer  # noqa: F401
from .fusion import fuse_level, up2_conv1x1_act, decode_fused, forward_sync_free, install  # noqa: F401
from .sharding import shard_clips, gather_outputs, PeerGather, row_band, search_transfer_rows, gather_rows  # noqa: F401
from .pipeline import HostPipeline  # noqa: F401
from .rl_deconv import create_blur_kernel, r_l_per_channel  # noqa: F401

__all__ = ["SearchTransfer", "SelfTransfer", "search_transfer", "fuse_level", "up2_conv1x1_act", "decode_fused", "forward_sync_free", "install",
           "shard_clips", "gather_outputs", "PeerGather", "HostPipeline", "create_blur_kernel", "r_l_per_channel", "load_library", "LIB_PATH"]
