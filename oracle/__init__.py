"""CPU oracle for the VAD inference hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain PyTorch/NumPy/SciPy (CPU) restatement of the reference
path  waveform -> lhotse-style fbank -> PyanNet2 / PyanNet(SincNet) forward ->
threshold + median filter -> run-length segments.  Every function cites the
reference file:line it follows (paths relative to /root/reference).

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- as the checker or
as the timed CPU baseline, never as part of the product path.  Nothing under
``universal-voice-activity-detection_b200/`` imports it.

PINNING.  The reference ships no tests, golden vectors or fixtures and its modules do not
import as they are in this image (pytorch_lightning, lhotse, asteroid_filterbanks, torchmetrics, ml_collections are
absent, no network).  What pins the oracle:
  * PINNED against outputs of the reference's OWN code run in this container (tests/golden/make_reference_golden.py:
    the reference modules are imported from /root/reference with inert stubs for the absent packages, which carry no
    arithmetic on these paths; fixtures tests/golden/reference_golden.{npz,json}; checked by
    tests/test_reference_golden.py): PyanNet2 construction + forward and VadModel.forward / predict_step (models.py),
    median_filter, the RLE + seconds arithmetic of both time bases, merge / split, the DER helpers (postproc.py), the
    frame arithmetic (receptive_field.py); SincNet / PyanNet structure (the reference's classes run with this package's
    ParamSincFB / Encoder as their filterbank).  Under manual_seed(42) the oracle's weights equal the reference's bit for bit.
  * PARITY UNPINNED for the two pieces whose arithmetic lives in third-party code that is not under /root/reference:
    lhotse ``Fbank`` (un-pinned editable checkout, requirements.txt:13; fbank.py) and asteroid-filterbanks==0.4
    ``ParamSincFB`` (requirements.txt:1; the filter synthesis inside models.SincNet -- the rest of SincNet is
    torch modules).  Those are restated from upstream knowledge and held by (a) being built from the same torch
    primitives (torch.fft.rfft, nn.Conv1d, nn.InstanceNorm1d, nn.MaxPool1d), (b) the structural constants the reference
    states (293 SincNet frames / 5 s, receptive field 991 / step 270, 500 fbank frames / 5 s) and (c) a bit-for-bit
    check of the Kaldi mel banks against torchaudio.
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
