"""Drop-in for ``DeepNetworks.HRNet.HRNet`` (reference src/DeepNetworks/HRNet.py:172-211).

Same constructor (``HRNet(config["network"])``), same ``forward(lrs, alphas)``
signature and result shape, same 31 ``state_dict`` keys/shapes (so
``load_state_dict(torch.load("HRNet.pth"))`` from predict.py:98-99 works), same
``.to(device)`` / ``.eval()`` / ``.parameters()`` behaviour -- but ``forward`` runs
the hand-written sm_100a kernels behind the C ABI (include/hrn_b200.h) instead of
ATen/cuDNN.  PyTorch is only used for parameter storage, device memory and streams.

Inference only: there is no autograd through this path and no CPU fallback.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib
from ._native import NativeHandleMixin


def _container(**children) -> nn.Module:
    m = nn.Module()
    for name, child in children.items():
        m.add_module(name, child)
    return m


def _conv_prelu_pair(c: int, k: int) -> nn.Sequential:
    # parameter holder for "conv, PReLU, conv, PReLU" (state_dict keys block.0 .. block.3)
    return nn.Sequential(nn.Conv2d(c, c, k, padding=k // 2), nn.PReLU(), nn.Conv2d(c, c, k, padding=k // 2), nn.PReLU())


class HRNet(NativeHandleMixin, nn.Module):
    """HRNet(config) with config = the "network" block of config/config.json."""

    def __init__(self, config):
        super().__init__()
        enc, rec, dec = config["encoder"], config["recursive"], config["decoder"]
        self._cfg = _lib.HrnConfig(
            enc_in_channels=enc["in_channels"], enc_num_layers=enc["num_layers"], enc_kernel_size=enc["kernel_size"],
            enc_channels=enc["channel_size"], rec_alpha_residual=int(bool(rec["alpha_residual"])),
            rec_in_channels=rec["in_channels"], rec_kernel_size=rec["kernel_size"],
            dec_in_channels=dec["deconv"]["in_channels"], dec_kernel_size=dec["deconv"]["kernel_size"],
            dec_stride=dec["deconv"]["stride"], dec_out_channels=dec["deconv"]["out_channels"],
            fin_in_channels=dec["final"]["in_channels"], fin_kernel_size=dec["final"]["kernel_size"],
            fin_out_channels=dec["final"]["out_channels"])
        c, k = enc["channel_size"], enc["kernel_size"]
        # Parameter containers only (never called): they give the reference's state_dict names and default init.
        self.encode = _container(
            init_layer=nn.Sequential(nn.Conv2d(enc["in_channels"], c, k, padding=k // 2), nn.PReLU()),
            res_layers=nn.Sequential(*[_container(block=_conv_prelu_pair(c, k)) for _ in range(enc["num_layers"])]),
            final=nn.Sequential(nn.Conv2d(c, c, k, padding=k // 2)))
        f, fk = rec["in_channels"], rec["kernel_size"]
        self.fuse = _container(fuse=nn.Sequential(
            _container(block=_conv_prelu_pair(2 * f, fk)), nn.Conv2d(2 * f, f, fk, padding=fk // 2), nn.PReLU()))
        d, fin = dec["deconv"], dec["final"]
        self.decode = _container(
            deconv=nn.Sequential(nn.ConvTranspose2d(d["in_channels"], d["out_channels"], d["kernel_size"],
                                                    stride=d["stride"]), nn.PReLU()),
            final=nn.Conv2d(fin["in_channels"], fin["out_channels"], fin["kernel_size"],
                            padding=fin["kernel_size"] // 2))
        self._native_init()       # device index -> [native handle address, weight fingerprint] (see _native.py)

    # ------------------------------------------------------------------ native handle
    def _handle_for(self, device: torch.device) -> ctypes.c_void_p:
        lib = _lib.load()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        entry = self._handles.get(idx)
        fp = self._fingerprint()
        if entry is None:
            handle = ctypes.c_void_p()
            _lib.check(lib.hrn_create(ctypes.byref(self._cfg), idx, ctypes.byref(handle)), "hrn_create")
            entry = [handle.value, None]
            self._handles[idx] = entry
        if entry[1] != fp:
            for key, tensor in self.state_dict().items():
                host = tensor.detach().to("cpu", torch.float32).contiguous()
                shape = (ctypes.c_int64 * host.dim())(*host.shape)
                _lib.check(lib.hrn_set_weight(ctypes.c_void_p(entry[0]), key.encode(), ctypes.c_void_p(host.data_ptr()),
                                              shape, host.dim()), f"hrn_set_weight({key})")
            entry[1] = fp
        return ctypes.c_void_p(entry[0])

    def __del__(self):
        try:
            lib = _lib.load()
            self._destroy_handles(lambda addr: lib.hrn_destroy(ctypes.c_void_p(addr)))
        except Exception:
            pass

    def reserve(self, device, b: int, l: int, h: int, w: int) -> None:
        """Sizes the native workspace for forwards of up to (b, l, h, w) now (hrn_reserve), so that no later forward has
        to reallocate it in the middle of the stream."""
        _lib.check(_lib.load().hrn_reserve(self._handle_for(torch.device(device)), int(b), int(l), int(h), int(w)),
                   "hrn_reserve")

    def debug_set(self, device, knob: str, value: int) -> None:
        """Bring-up knobs of the native handle (see hrn_debug_set in include/hrn_b200.h)."""
        handle = self._handle_for(torch.device(device))
        _lib.check(_lib.load().hrn_debug_set(handle, knob.encode(), int(value)), "hrn_debug_set")

    PROFILE_CLASSES = ("conv3x3_umma<64>", "conv3x3_umma<128>", "conv_init", "decoder", "median_anchor",
                       "resblock64_umma", "live_lists", "forward_span", "fuse_wave", "enc_wave")

    def profile_begin(self, device) -> None:
        """Arm per-launch CUDA-event timing of the following forward calls (hrn_profile_begin)."""
        _lib.check(_lib.load().hrn_profile_begin(self._handle_for(torch.device(device))), "hrn_profile_begin")

    def profile_end(self, device) -> dict:
        """-> {kernel class: {"ms": total, "flops": total algorithmic, "launches": n}} (hrn_profile_end)."""
        n = len(self.PROFILE_CLASSES)
        ms, fl, cnt = (ctypes.c_double * n)(), (ctypes.c_double * n)(), (ctypes.c_int64 * n)()
        _lib.check(_lib.load().hrn_profile_end(self._handle_for(torch.device(device)), ms, fl, cnt), "hrn_profile_end")
        return {name: {"ms": ms[i], "flops": fl[i], "launches": int(cnt[i])}
                for i, name in enumerate(self.PROFILE_CLASSES)}

    # ------------------------------------------------------------------ forward
    def _check_inputs(self, lrs, alphas):
        _lib.require_cuda_tensor(lrs, "lrs")
        _lib.require_cuda_tensor(alphas, "alphas")
        if lrs.dim() != 4:
            raise ValueError("lrs must be (B, L, H, W)")
        b, l, h, w = lrs.shape
        if h != w:
            raise ValueError("square inputs only: the reference view() at HRNet.py:204 swaps H and W otherwise")
        if alphas.numel() != b * l:
            raise ValueError("alphas must be (B, L)")
        if self.training and torch.is_grad_enabled():
            raise RuntimeError("the B200 HRNet path is inference-only; call .eval() (autograd is not supported)")
        lrs = lrs.detach().to(torch.float32).contiguous()
        alphas = alphas.detach().to(device=lrs.device, dtype=torch.float32).contiguous()
        return lrs, alphas, b, l, h, w

    def forward(self, lrs, alphas):
        """lrs (B, L, H, W) in [0, 1], alphas (B, L) in {0, 1} -> srs (B, 1, 3H, 3W) float32 on lrs.device."""
        lrs, alphas, b, l, h, w = self._check_inputs(lrs, alphas)
        handle = self._handle_for(lrs.device)
        stride = self._cfg.dec_stride
        srs = torch.empty((b, self._cfg.fin_out_channels, stride * h, stride * w), dtype=torch.float32,
                          device=lrs.device)
        with torch.cuda.device(lrs.device):
            _lib.check(_lib.load().hrn_forward(handle, lrs.data_ptr(), alphas.data_ptr(), b, l, h, w, srs.data_ptr(),
                                               _lib.current_stream_ptr(lrs.device)), "hrn_forward")
        return srs

    def forward_host(self, lrs_host: torch.Tensor, alphas_host: torch.Tensor, out_host: torch.Tensor = None,
                     device="cuda:0") -> torch.Tensor:
        """Host-buffer variant of the train.py:200-208 pattern (H2D, forward, D2H) through hrn_forward_host.
        Inputs are CPU float32 tensors (pinned memory recommended); returns a CPU tensor.  ``lrs_host`` may also be a
        uint16 tensor holding the raw 16-bit views (DataLoader.py:134): it is then copied as is and scaled to [0, 1] on
        the device exactly like DataLoader.py:195-198 (hrn_forward_host_u16)."""
        if lrs_host.is_cuda or alphas_host.is_cuda:
            raise ValueError("forward_host takes host tensors")
        raw16 = lrs_host.dtype == torch.uint16
        lrs_host = lrs_host.contiguous() if raw16 else lrs_host.to(torch.float32).contiguous()
        alphas_host = alphas_host.to(torch.float32).contiguous()
        b, l, h, w = lrs_host.shape
        device = torch.device(device)
        handle = self._handle_for(device)
        if out_host is None:
            out_host = torch.empty((b, 1, 3 * h, 3 * w), dtype=torch.float32, pin_memory=True)
        with torch.cuda.device(device):
            entry = _lib.load().hrn_forward_host_u16 if raw16 else _lib.load().hrn_forward_host
            _lib.check(entry(handle, lrs_host.data_ptr(), alphas_host.data_ptr(), b, l, h, w, out_host.data_ptr(),
                             _lib.current_stream_ptr(device)), "hrn_forward_host_u16" if raw16 else "hrn_forward_host")
        return out_host

    def forward_host_submit(self, lrs_host: torch.Tensor, alphas_host: torch.Tensor, out_host: torch.Tensor = None,
                            device="cuda:0"):
        """Asynchronous forward_host (hrn_forward_host_submit): enqueues H2D, forward and D2H and returns a pending
        handle for forward_host_wait.  Two calls may be in flight, so the copies of one batch overlap the kernels of
        its neighbours; use pinned float32 tensors and do not touch them (or ``out_host``) before the wait."""
        if lrs_host.is_cuda or alphas_host.is_cuda:
            raise ValueError("forward_host_submit takes host tensors")
        lrs_host = lrs_host.to(torch.float32).contiguous()
        alphas_host = alphas_host.to(torch.float32).contiguous()
        b, l, h, w = lrs_host.shape
        device = torch.device(device)
        handle = self._handle_for(device)
        if out_host is None:
            out_host = torch.empty((b, 1, 3 * h, 3 * w), dtype=torch.float32, pin_memory=True)
        ticket = ctypes.c_int64(0)
        with torch.cuda.device(device):
            _lib.check(_lib.load().hrn_forward_host_submit(handle, lrs_host.data_ptr(), alphas_host.data_ptr(), b, l, h, w,
                                                           out_host.data_ptr(), _lib.current_stream_ptr(device),
                                                           ctypes.byref(ticket)), "hrn_forward_host_submit")
        return (handle, ticket.value, out_host, lrs_host, alphas_host)     # the host tensors stay alive until the wait

    def forward_host_wait(self, pending) -> torch.Tensor:
        """Blocks until the call behind ``pending`` (from forward_host_submit) is complete; returns its SR host tensor."""
        handle, ticket, out_host = pending[0], pending[1], pending[2]
        _lib.check(_lib.load().hrn_forward_host_wait(handle, ticket), "hrn_forward_host_wait")
        return out_host

    def forward_host_iter(self, batches, device="cuda:0"):
        """Pipelined validation loop (train.py:199-208): yields the SR host tensor of every (lrs_host, alphas_host)
        batch in order while the next batch is already copying / computing."""
        pending = None
        for lrs_host, alphas_host in batches:
            nxt = self.forward_host_submit(lrs_host, alphas_host, device=device)
            if pending is not None:
                yield self.forward_host_wait(pending)
            pending = nxt
        if pending is not None:
            yield self.forward_host_wait(pending)

    def forward_stage(self, lrs, alphas, stage: int, shape):
        """Test hook: run forward and return (srs, fp32 NCHW copy of the named intermediate)."""
        lrs, alphas, b, l, h, w = self._check_inputs(lrs, alphas)
        handle = self._handle_for(lrs.device)
        srs = torch.empty((b, 1, 3 * h, 3 * w), dtype=torch.float32, device=lrs.device)
        dump = torch.empty(tuple(shape), dtype=torch.float32, device=lrs.device)
        with torch.cuda.device(lrs.device):
            _lib.check(_lib.load().hrn_forward_dump(handle, lrs.data_ptr(), alphas.data_ptr(), b, l, h, w,
                                                    srs.data_ptr(), int(stage), dump.data_ptr(),
                                                    _lib.current_stream_ptr(lrs.device)), "hrn_forward_dump")
        return srs, dump


def stage_anchor() -> int:
    return 1


def stage_enc(i: int) -> int:
    return 0x100 + i


def stage_fuse(level: int, j: int) -> int:
    return 0x200 + 4 * level + j
