"""Drop-in for src/utils/receptive_field.py (conv arithmetic of the SincNet front-end):
conv1d_num_frames :28-55, multi_conv_num_frames :58-71, get_num_frames :165-193,
receptive_field_size :196-219 of the reference.  Pure integer host code."""

from functools import lru_cache
from typing import List

from b200vad.host import SINC_KERNELS, SINC_STRIDES, conv1d_num_frames  # noqa: F401


def multi_conv_num_frames(num_samples: int, kernel_size: List[int] = None, stride: List[int] = None,
                          padding: List[int] = None, dilation: List[int] = None) -> int:
    num_frames = num_samples
    for k, s, p, d in zip(kernel_size, stride, padding, dilation):
        num_frames = conv1d_num_frames(num_frames, kernel_size=k, stride=s, padding=p, dilation=d)
    return num_frames


def conv1d_receptive_field_size(num_frames=1, kernel_size=5, stride=1, dilation=1):
    return 1 + (kernel_size - 1) * dilation + (num_frames - 1) * stride


def multi_conv_receptive_field_size(num_frames: int, kernel_size=None, stride=None, padding=None, dilation=None) -> int:
    size = num_frames
    for k, s, d in reversed(list(zip(kernel_size, stride, dilation))):
        size = conv1d_receptive_field_size(num_frames=size, kernel_size=k, stride=s, dilation=d)
    return size


def conv1d_receptive_field_center(frame=0, kernel_size=5, stride=1, padding=0, dilation=1) -> int:
    return frame * stride + ((kernel_size - 1) * dilation) // 2 - padding


def multi_conv_receptive_field_center(frame: int, kernel_size=None, stride=None, padding=None, dilation=None) -> int:
    center = frame
    for k, s, p, d in reversed(list(zip(kernel_size, stride, padding, dilation))):
        center = conv1d_receptive_field_center(frame=center, kernel_size=k, stride=s, padding=p, dilation=d)
    return center


@lru_cache
def get_num_frames(num_samples: int) -> int:
    n = len(SINC_KERNELS)
    return int(multi_conv_num_frames(num_samples, kernel_size=list(SINC_KERNELS), stride=list(SINC_STRIDES),
                                     padding=[0] * n, dilation=[1] * n))


def receptive_field_size(num_frames: int = 1) -> int:
    n = len(SINC_KERNELS)
    return multi_conv_receptive_field_size(num_frames, kernel_size=list(SINC_KERNELS), stride=list(SINC_STRIDES),
                                           dilation=[1] * n)
