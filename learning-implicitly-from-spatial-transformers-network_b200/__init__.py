"""list_b200 -- B200-native drop-in for LIST's per-query SDF hot path.

Mirrors the reference's module layout for that path only
(`network.modules.PerceptualPooling` / `VoxelDecoder2`, `network.models.LIST`,
`network.executors.LIST`) on top of a C-ABI CUDA library (`csrc/`,
`include/list_b200.h`).  There is no CPU fallback: every compute entry point
raises if `liblist_b200.so` is missing or no B200 is visible.
"""
__version__ = "0.1.0"
