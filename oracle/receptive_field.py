"""Oracle: SincNet conv arithmetic (frames per sample count, receptive field).

TEST INFRASTRUCTURE ONLY.  Follows src/utils/receptive_field.py:28-55 (conv1d_num_frames),
:58-71 (multi-conv fold), :165-193 (get_num_frames), :196-219 (receptive_field_size).
Constants RF=991 / STEP=270 / HALF=495 are the ones duplicated at
src/datasets/custom_vad.py:43-45 and src/scripts/predict_sincnet.py:494-496.
"""

KERNELS = (251, 3, 5, 3, 5, 3)
STRIDES = (10, 3, 1, 3, 1, 3)


def conv1d_num_frames(num_samples, kernel_size=5, stride=1, padding=0, dilation=1):
    return 1 + (num_samples + 2 * padding - dilation * (kernel_size - 1) - 1) // stride


def get_num_frames(num_samples):
    n = num_samples
    for k, s in zip(KERNELS, STRIDES):
        n = conv1d_num_frames(n, kernel_size=k, stride=s)
    return int(n)


def receptive_field_size(num_frames=1):
    size = num_frames
    for k, s in reversed(list(zip(KERNELS, STRIDES))):
        size = k + (size - 1) * s
    return size
