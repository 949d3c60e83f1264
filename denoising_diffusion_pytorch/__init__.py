"""Drop-in import path of the reference (`main.py:6`): re-exports the B200-native implementations."""
from videometamaterials_b200 import GaussianDiffusion, Trainer, Unet3D  # noqa: F401
