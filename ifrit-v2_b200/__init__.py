"""ifrit-v2_b200 — B200-native tile rasterizer behind Ifrit-v2's soft-renderer API.

Contents (only what the hot path needs):
  csrc/       sm_100a CUDA kernels + the C-ABI (include/ifrit_b200.h) -> lib/libifrit_b200.so
  renderer.py host-side mirror of the reference's TileRasterRendererCuda class over the C-ABI
  scenes.py   seeded synthetic scenes of SURVEY.md §8(d)
  multigpu.py sort-first tile partition + NCCL gather plumbing (torch.distributed)
"""
from . import scenes  # noqa: F401
