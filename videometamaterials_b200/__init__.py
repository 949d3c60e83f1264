"""B200-native (sm_100a) implementation of the VideoMetamaterials video-diffusion hot path.

Public surface = the reference's (`from denoising_diffusion_pytorch import Unet3D, GaussianDiffusion, Trainer`).
"""
from .unet3d import Unet3D
from .diffusion import GaussianDiffusion

__all__ = ["Unet3D", "GaussianDiffusion"]
