"""CPU oracle for the VAD inference hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain PyTorch/NumPy/SciPy (CPU) restatement of the reference
path  waveform -> lhotse-style fbank -> PyanNet2 / PyanNet(SincNet) forward ->
threshold + median filter -> run-length segments.  Every function cites the
reference file:line it follows (paths relative to /root/reference).

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- as the checker or
as the timed CPU baseline, never as part of the product path.  Nothing under
``universal-voice-activity-detection_b200/`` imports it.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
 and its own modules cannot be imported in this image
(pytorch_lightning, lhotse, asteroid_filterbanks, torchmetrics, ml_collections
are absent, no network).  The oracle is therefore pinned only by (a) being built
from the very same torch / scipy primitives the reference calls (nn.LSTM,
nn.Linear, nn.Conv1d, nn.InstanceNorm1d, nn.MaxPool1d, torch.fft.rfft,
scipy.signal.medfilt), (b) the structural constants the reference states
(293 SincNet frames / 5 s, receptive field 991 / step 270, 500 fbank frames
/ 5 s), and (c) cross-checks against torchaudio's Kaldi mel banks.  Third-party
arithmetic restated from upstream knowledge: lhotse ``Fbank`` (un-pinned editable
checkout, requirements.txt:13) and asteroid-filterbanks==0.4 ``ParamSincFB``
(requirements.txt:1).
"""

from .fbank import lhotse_fbank, kaldi_mel_banks, povey_window, num_fbank_frames  # noqa: F401
from .models import SincNet, PyanNet, PyanNet2, VadModel, ParamSincFB  # noqa: F401
from .postproc import (  # noqa: F401
    median_filter,
    median_window,
    rle_segments,
    rle_segments_sincnet,
    merge_intervals_with_buffer,
    split_into_windows,
    slice_recordings,
    get_binary_tensor,
    get_false_alarm,
    get_missed_detection,
)
from .receptive_field import get_num_frames, conv1d_num_frames, receptive_field_size  # noqa: F401
