"""The fused frame-staging op, exposed where the reference keeps its transforms
(src/transforms/; the reference's only file there is an unused MNIST ToTensor()).
`GrayStack()(frames_u8)` = RGB u8 (n,H,W,3) -> (n-frame_skip+... ) see below; one CUDA kernel."""
import torch

from carla_imitation_learning_b200.engine import sliding_window, stage_gray


class GrayStack:
    """frames (n,H,W,3) u8 on the device -> (n-frame_skip, frame_skip, H, W) window view of the
    gray planes (0.299R+0.587G+0.114B)/255, as imitation_dataset.py:115-133 builds per sample."""

    def __init__(self, frame_skip: int = 4, dtype: torch.dtype = torch.float32):
        self.frame_skip, self.dtype = frame_skip, dtype

    def __call__(self, frames_u8: torch.Tensor) -> torch.Tensor:
        return sliding_window(stage_gray(frames_u8, dtype=self.dtype), self.frame_skip)
