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
    """Lazily (re)built kernel-layout copy of a module's fp32 parameters.

    The copy is rebuilt when a parameter / buffer was replaced or written in place through autograd-visible ops (the key is
    ``(device, data_ptr, _version)`` per tensor), after ``load_state_dict`` and after ``.to()`` / ``.cuda()`` / ``.half()``
    (hooks in the owning module call :meth:`invalidate`).  Writes that bypass the version counter -- ``p.data.copy_(ema)``,
    ``p.data.mul_()``, numpy views, shared storage -- are NOT visible to that key: call ``module.invalidate_packed()`` after
    them, or set ``module.repack_always = True`` to rebuild the copy on every forward (about 30 tiny launches), which is
    what the reference's "read the live weights on every forward" amounts to.
    """

    def __init__(self):
        self.blob = None
        self.key = None

    def invalidate(self):
        self.blob = None
        self.key = None

    def get(self, module: nn.Module, device, build, always: bool = False):
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in module.parameters()) + \
            tuple((b.data_ptr(), b._version) for b in module.buffers())
        if always or self.blob is None or key != self.key:
            self.blob = build()
            self.key = key
        return self.blob


class PackedOwner:
    """Mixin for modules that own PackedWeights: invalidation hooks + the explicit knobs documented above."""

    repack_always = False

    def _packed_caches(self):
        return [v for v in self.__dict__.values() if isinstance(v, PackedWeights)]

    def invalidate_packed(self):
        for c in self._packed_caches():
            c.invalidate()
        for m in self.children():
            if isinstance(m, PackedOwner):
                m.invalidate_packed()

    def _load_from_state_dict(self, *args, **kwargs):
        self.invalidate_packed()
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self.invalidate_packed()
        return super()._apply(fn, *args, **kwargs)
