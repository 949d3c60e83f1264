"""B200-native (sm_100a) implementation of the VideoMetamaterials video-diffusion hot path.

Public surface = the reference's (`from denoising_diffusion_pytorch import Unet3D, GaussianDiffusion, Trainer`)
plus an Accelerate-compatible `Accelerator` over NCCL.
"""
from .unet3d import Unet3D
from .diffusion import GaussianDiffusion
from .trainer import Trainer
from .accel import Accelerator, DistributedDataParallelKwargs, InitProcessGroupKwargs, broadcast_object_list

__all__ = ["Unet3D", "GaussianDiffusion", "Trainer", "Accelerator", "DistributedDataParallelKwargs", "InitProcessGroupKwargs",
           "broadcast_object_list"]
