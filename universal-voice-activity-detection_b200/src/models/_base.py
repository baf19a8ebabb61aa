"""Base class shim: ``pl.LightningModule`` when Lightning is importable, else ``nn.Module`` with the
two Lightning features the reference models use (``save_hyperparameters`` / ``hparams``)."""

import inspect

import torch.nn as nn

try:  # pragma: no cover - Lightning is absent in the build image
    import pytorch_lightning as pl

    Base = pl.LightningModule
except Exception:  # noqa: BLE001

    class _HParams(dict):
        __getattr__ = dict.__getitem__

        def __setattr__(self, k, v):
            self[k] = v

    class Base(nn.Module):
        def save_hyperparameters(self, *names):
            frame = inspect.currentframe().f_back
            if "hparams" not in self.__dict__:
                self.__dict__["hparams"] = _HParams()
            for n in names:
                self.hparams[n] = frame.f_locals[n]


class PackedWeights:
    """Lazily (re)built kernel-layout copy of a module's fp32 parameters."""

    def __init__(self):
        self.blob = None
        self.key = None

    def get(self, module: nn.Module, device, build):
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in module.parameters()) + \
            tuple((b.data_ptr(), b._version) for b in module.buffers())
        if self.blob is None or key != self.key:
            self.blob = build()
            self.key = key
        return self.blob
