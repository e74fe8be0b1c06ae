"""CPU suite for the boundary: the C-ABI library builds/loads, exports every symbol that
include/hrn_b200.h declares, fails loudly without a GPU (no fallback), and the Python host
mirror keeps the reference's module surface (state_dict keys, signatures).  No compute
calls are made here."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from highres_net_b200 import _lib
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hrn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hrn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from highres_net_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/hrn_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared, "ctypes binding table out of sync with the header"
    assert lib.hrn_abi_version() == 1
    assert lib.hrn_kernel_launch_count() == 0


def test_library_is_in_tree_and_built_for_sm100a(lib):
    import shutil
    import subprocess
    import highres_net_b200 as hb
    path = hb.library_path()
    assert path.startswith(ROOT) and os.path.exists(path)                 # in-tree, not site-packages
    build_py = open(os.path.join(ROOT, "highres-net_b200", "build.py")).read()
    assert "arch=compute_100a,code=sm_100a" in build_py and "-lineinfo" in build_py
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if os.path.exists(cuobjdump):
        elfs = subprocess.run([cuobjdump, "-lelf", path], capture_output=True, text=True).stdout
        assert "sm_100a" in elfs and "sm_90" not in elfs                  # one target only: no multi-arch fat binary


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu(lib):
    from highres_net_b200 import _lib
    cfg = _lib.HrnConfig(2, 2, 3, 64, 1, 64, 3, 64, 3, 3, 64, 64, 1, 1)
    handle = ctypes.c_void_p()
    rc = lib.hrn_create(ctypes.byref(cfg), 0, ctypes.byref(handle))
    assert rc != 0 and not handle
    assert "no CPU fallback" in _lib.last_error() or "CUDA" in _lib.last_error()


def test_create_rejects_unsupported_config(lib):
    from highres_net_b200 import _lib
    cfg = _lib.HrnConfig(2, 2, 5, 64, 1, 64, 3, 64, 3, 3, 64, 64, 1, 1)      # kernel_size 5
    handle = ctypes.c_void_p()
    assert lib.hrn_create(ctypes.byref(cfg), 0, ctypes.byref(handle)) != 0
    assert "unsupported network config" in _lib.last_error()


def test_hrnet_module_surface_matches_reference():
    import highres_net_b200 as hb
    from oracle import hrnet_oracle
    net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG)
    sd = net.state_dict()
    shapes = hrnet_oracle.param_shapes()
    assert list(sd) == list(shapes)                                   # same 31 keys, same order
    assert all(tuple(sd[k].shape) == shapes[k] for k in shapes)
    assert sum(p.numel() for p in net.parameters()) == 591818
    net.load_state_dict(hrnet_oracle.make_params(0), strict=True)
    assert list(inspect.signature(net.forward).parameters) == ["lrs", "alphas"]
    assert net.training and not net.eval().training


def test_host_functions_keep_reference_signatures():
    import highres_net_b200 as hb
    sig = inspect.signature(hb.lanczos_shift)
    assert list(sig.parameters) == ["img", "shift", "p", "a", "N"]
    assert [sig.parameters[k].default for k in ("p", "a", "N")] == [3, 3, 7]          # lanczos.py:47
    sig = inspect.signature(hb.lanczos_kernel)
    assert list(sig.parameters) == ["dx", "a", "N", "dtype", "device"]                # lanczos.py:5
    sig = inspect.signature(hb.shift_cPSNR)
    assert list(sig.parameters)[:4] == ["sr", "hr", "hr_map", "border_w"] and sig.parameters["border_w"].default == 3
    assert list(inspect.signature(hb.cPSNR).parameters) == ["sr", "hr", "hr_map"]     # Evaluator.py:11


def test_no_cpu_fallback_in_host_layer():
    import highres_net_b200 as hb
    from oracle import hrnet_oracle
    net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.rand(1, 2, 8, 8), torch.ones(1, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hb.lanczos_shift(torch.rand(1, 1, 8, 8), torch.zeros(1, 2))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under highres-net_b200/ may reference it."""
    pkg = os.path.join(ROOT, "highres-net_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "/root/reference" not in text, f


def test_module_copies_and_pickles_without_its_native_handles():
    """The reference module supports copy.deepcopy / torch.save(model) (best-model snapshots, EMA copies); the native
    handles are process-local pointers and must stay out of the copied state (a copy makes its own lazily)."""
    import copy
    import io
    import highres_net_b200 as hb
    from oracle import hrnet_oracle
    for net in (hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG), hb.ShiftNet()):
        net._handles[0] = [0xdeadbeef, ("fingerprint",)]          # what a first forward on cuda:0 would have stored
        twin = copy.deepcopy(net)
        assert twin._handles == {} and net._handles[0][0] == 0xdeadbeef
        assert all(torch.equal(a, b) for a, b in zip(twin.state_dict().values(), net.state_dict().values()))
        buf = io.BytesIO()
        torch.save(net, buf)
        buf.seek(0)
        loaded = torch.load(buf, weights_only=False)
        assert loaded._handles == {} and list(loaded.state_dict()) == list(net.state_dict())
        net._handles.clear()                                      # nothing real to destroy


def test_weight_refresh_rule_sees_what_torch_can_and_cannot_signal():
    """`.data` edits bump neither data_ptr nor _version: invalidate_weights() or verify_weights cover them."""
    import highres_net_b200 as hb
    from oracle import hrnet_oracle
    net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG)
    fp0 = net._fingerprint()
    with torch.no_grad():
        net.encode.final[0].bias.add_(1.0)                        # version bump: seen
    fp1 = net._fingerprint()
    assert fp1 != fp0
    net.encode.final[0].bias.data.mul_(2.0)                       # .data edit: invisible to the cheap rule ...
    assert net._fingerprint() == fp1
    net.invalidate_weights()                                      # ... so the caller says so,
    fp2 = net._fingerprint()
    assert fp2 != fp1
    net.verify_weights = True                                     # or asks for the checksum rule
    fp3 = net._fingerprint()
    net.encode.final[0].bias.data.mul_(0.5)
    assert net._fingerprint() != fp3
    net.load_state_dict(hrnet_oracle.make_params(0))              # wholesale rewrites invalidate by themselves
    assert net._fingerprint()[0] > fp3[0]
