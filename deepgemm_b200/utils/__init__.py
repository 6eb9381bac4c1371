from .math import (  # noqa: F401
    align, ceil_div, ceil_to_ue8m0, pack_ue8m0_to_int, unpack_ue8m0_from_int,
    per_block_cast_to_fp8, per_channel_cast_to_fp8, per_custom_dims_cast_to_fp8, per_token_cast_to_fp8,
)
from .layout import *  # noqa: F401,F403
