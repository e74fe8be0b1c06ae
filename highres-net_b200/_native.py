"""Book-keeping shared by the nn.Module drop-ins (HRNet, ShiftNet): one native handle per device, kept OUT of the
module's copied / pickled state, and the rule that decides when the native weights must be uploaded again."""
from __future__ import annotations

import torch


class NativeHandleMixin:
    """The torch parameters are the master copy; the native handle holds a repacked bf16 copy.  The copy is refreshed
    when (a) a parameter was replaced or modified through autograd-visible ops (``data_ptr`` / ``_version`` change:
    ``load_state_dict``, ``.to()``, optimiser steps, ``copy_`` under ``no_grad``), (b) ``invalidate_weights()`` was
    called, or (c) ``verify_weights`` is on and a checksum of the parameter values changed.

    In-place edits through ``.data`` (``p.data.mul_(2)``, manual EMA, ``fc2.weight.data.zero_()``) bump NEITHER the
    pointer NOR the version counter, so PyTorch gives no cheap signal for them: call ``invalidate_weights()`` after
    such an edit, or set ``module.verify_weights = True`` (one small reduction and a host sync per forward)."""

    verify_weights = False

    def _native_init(self):
        # plain dict of plain ints: device index -> [handle address, fingerprint]; dropped by __getstate__
        object.__setattr__(self, "_handles", {})
        object.__setattr__(self, "_weights_epoch", 0)

    def invalidate_weights(self) -> None:
        """Forces the next forward to upload the current parameter values to every native handle."""
        object.__setattr__(self, "_weights_epoch", self._weights_epoch + 1)

    def _native_tensors(self):
        return list(self.parameters()) + list(self.buffers())

    def _fingerprint(self):
        tensors = self._native_tensors()
        fp = (self._weights_epoch,) + tuple((t.data_ptr(), t._version) for t in tensors)
        if self.verify_weights:
            with torch.no_grad():
                flat = torch.cat([t.detach().reshape(-1).to(torch.float64) for t in tensors if t.is_floating_point()])
                idx = torch.arange(1, flat.numel() + 1, dtype=torch.float64, device=flat.device)
                fp += (float(flat.sum()), float((flat * idx).sum()))
        return fp

    # ---- nn.Module hooks: anything that rewrites parameters wholesale also invalidates explicitly
    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_weights()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate_weights()
        return out

    # ---- copy / pickle: the handles are process-local native pointers; a copy creates its own lazily
    def __getstate__(self):
        state = dict(self.__dict__)
        state["_handles"] = {}
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        object.__setattr__(self, "_handles", {})

    def _destroy_handles(self, destroy_fn) -> None:
        handles, self.__dict__["_handles"] = self.__dict__.get("_handles", {}), {}
        for addr, _ in handles.values():
            try:
                destroy_fn(addr)
            except Exception:
                pass
