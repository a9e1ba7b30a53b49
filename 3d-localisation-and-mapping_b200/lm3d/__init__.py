"""lm3d -- B200-native 2D-box -> 3D lift (host side of ``liblm3d.so``).

``lm3d.lift``  torch-facing wrappers over the C ABI (``include/lm3d.h``)
``lm3d.dist``  frame sharding across ranks + all-gather of the per-box records
``lm3d.synth`` seeded synthetic scan sequences
"""
__all__ = ["lift", "dist", "synth", "_capi"]
