"""Drop-in for ``DeepNetworks.ShiftNet.ShiftNet`` (src/DeepNetworks/ShiftNet.py) as an INFERENCE module, and for
``train.register_batch`` (src/train.py:26-44) -- row N3 of SURVEY.md section 8f.

Same constructor, same ``state_dict`` keys and shapes (a checkpoint saved from the reference loads with
``load_state_dict``), same ``forward(x) -> (N, 2)`` and ``transform(theta, I)``.  The forward runs in eval mode only:
BatchNorm uses its running statistics (folded into the convolutions), Dropout is the identity; everything numerical runs
in the CUDA library (hrn_shiftnet_forward: tcgen05 convolutions, a split-K tcgen05 GEMM for fc1).  The reference only
ever calls ShiftNet inside the training step (train-mode batch statistics, autograd): that use stays out of scope and
raises here."""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib
from ._native import NativeHandleMixin
from . import lanczos as _lanczos


def _block(cin, cout, pool):
    layers = [nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU()]
    if pool:
        layers.append(nn.MaxPool2d(2))
    return nn.Sequential(*layers)


class ShiftNet(NativeHandleMixin, nn.Module):
    """ShiftNet.py:6-75; the torch layers only hold the parameters and buffers (and define the state_dict layout)."""

    def __init__(self, in_channel=1):
        super().__init__()
        if in_channel != 1:
            raise ValueError("the B200 ShiftNet path supports in_channel=1 (pairs of single-channel crops) only")
        self.layer1 = _block(2 * in_channel, 64, False)
        self.layer2 = _block(64, 64, True)
        self.layer3 = _block(64, 64, False)
        self.layer4 = _block(64, 64, True)
        self.layer5 = _block(64, 128, False)
        self.layer6 = _block(128, 128, True)
        self.layer7 = _block(128, 128, False)
        self.layer8 = _block(128, 128, False)
        self.drop1 = nn.Dropout(p=0.5)
        self.fc1 = nn.Linear(128 * 16 * 16, 1024)
        self.activ1 = nn.ReLU()
        self.fc2 = nn.Linear(1024, 2, bias=False)
        self._native_init()
        self.fc2.weight.data.zero_()                 # ShiftNet.py:48: starts as the identity transformation

    def _handle_for(self, device):
        lib = _lib.load()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        entry = self._handles.get(idx)
        fp = self._fingerprint()
        if entry is None:
            handle = ctypes.c_void_p()
            _lib.check(lib.hrn_shiftnet_create(idx, ctypes.byref(handle)), "hrn_shiftnet_create")
            entry = [handle.value, None]
            self._handles[idx] = entry
        if entry[1] != fp:
            for key, tensor in self.state_dict().items():
                if key.endswith("num_batches_tracked"):
                    continue
                host = tensor.detach().to("cpu", torch.float32).contiguous()
                shape = (ctypes.c_int64 * host.dim())(*host.shape)
                _lib.check(lib.hrn_shiftnet_set_weight(ctypes.c_void_p(entry[0]), key.encode(),
                                                       ctypes.c_void_p(host.data_ptr()), shape, host.dim()),
                           f"hrn_shiftnet_set_weight({key})")
            entry[1] = fp
        return ctypes.c_void_p(entry[0])

    def __del__(self):
        try:
            lib = _lib.load()
            self._destroy_handles(lambda addr: lib.hrn_shiftnet_destroy(ctypes.c_void_p(addr)))
        except Exception:
            pass

    def debug_set(self, device, knob: str, value: int) -> None:
        """Bring-up knobs of the native handle (hrn_shiftnet_debug_set)."""
        _lib.check(_lib.load().hrn_shiftnet_debug_set(self._handle_for(torch.device(device)), knob.encode(), int(value)),
                   "hrn_shiftnet_debug_set")

    def forward(self, x):
        """x (N, 2, 128, 128): pairs cat([reference, view], 1) -> (N, 2) translation parameters (dx, dy)."""
        _lib.require_cuda_tensor(x, "x")
        if self.training:
            raise RuntimeError("the B200 ShiftNet path is eval-mode only (BatchNorm running statistics, no dropout, no "
                               "autograd); call .eval()")
        if x.dim() != 4 or x.shape[1] != 2:
            raise ValueError("x must be (N, 2, H, W)")
        x = x.detach().to(torch.float32).contiguous()
        n, _, h, w = x.shape
        handle = self._handle_for(x.device)
        theta = torch.empty((n, 2), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().hrn_shiftnet_forward(handle, x.data_ptr(), n, h, w, theta.data_ptr(),
                                                        _lib.current_stream_ptr(x.device)), "hrn_shiftnet_forward")
        return theta

    def transform(self, theta, I, device="cpu"):
        """ShiftNet.py:77-90: shift the images I (B, 1, H, W) by theta (B, 2) = (dx, dy) with the Lanczos kernel."""
        self.theta = theta
        return _lanczos.transform(theta, I)


def register_batch(shiftNet, lrs, reference):
    """train.py:26-44: thetas (B, V, 2) of every view of ``lrs`` (B, V, H, W) against ``reference`` (B, 1, H, W).
    The reference loops over the views (one forward per view); here all B * V pairs go through one forward."""
    b, v, h, w = lrs.shape
    pairs = torch.stack([reference.expand(b, v, h, w), lrs], dim=2).reshape(b * v, 2, h, w)
    return shiftNet(pairs).view(b, v, 2)
