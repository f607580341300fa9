"""Import shim for the UPSTREAM Point2Cyl reference (test infrastructure only).

The reference lives at /root/reference in the authoring container and does NOT
travel to the GPU box.  This module is therefore only used by
tests/golden/make_golden.py (run here, outputs committed) and by the optional
`-m "not gpu"` tests that re-validate the oracle restatement when the reference
tree happens to be present.  Nothing in the product path, the `-m gpu` tests,
smoke() or bench.py imports it.

What it patches:
  * stub modules for dependencies the reference imports at module scope but never
    uses on the forward+loss path: chamferdist, h5py, trimesh, plyfile, skimage
    (data_utils.py:5-24, losses.py:14, utils.py:7-11); torchgeometry gets the oracle's
    restatement of angle_axis_to_rotation_matrix (used by the projection functions)
  * torch.symeig (data_utils.py:170) was removed from torch; symeig defaulted to
    upper=True and returned ascending eigenvalues, which is
    torch.linalg.eigh(A, UPLO='U').
"""
import importlib
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("P2C_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "pointnet_util.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    class _ChamferDistance:  # losses.py:14-15 instantiates it at import time
        def __call__(self, *a, **k):
            raise RuntimeError("chamferdist stub: dead code on the hot path")

    _stub("chamferdist", ChamferDistance=_ChamferDistance)
    _stub("h5py")
    _stub("trimesh")
    # torchgeometry==0.1.2 cannot be installed offline; its one function on our path is restated in the oracle
    # (PARITY UNPINNED for that function, see oracle/p2c_oracle.py) and handed to the reference as `tgm`.
    from oracle import p2c_oracle as _orc
    _stub("torchgeometry", angle_axis_to_rotation_matrix=_orc.angle_axis_to_rotation_matrix)
    _stub("plyfile")
    sk = _stub("skimage")
    sk.measure = _stub("skimage.measure")
    # torch >= 1.13 keeps a `torch.symeig` that only raises; replace it either way.
    def symeig(A, eigenvectors=False, upper=True):
        e, v = torch.linalg.eigh(A, UPLO="U" if upper else "L")
        return e, v
    torch.symeig = symeig


_loaded = {}


def load():
    """Return a namespace with the reference modules: .util, .net, .losses, .data_utils."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REF_ROOT}")
    _install_stubs()
    # The reference resolves its own modules through sys.path (train_*.py:14-16).
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in
                  ("models", "models.pointnet_util", "models.pointnet_extrusion",
                   "losses", "data_utils", "utils", "global_variables", "pointnet_util")}
    for k in saved_mods:
        sys.modules.pop(k, None)
    sys.path.insert(0, os.path.join(REF_ROOT, "models"))
    sys.path.insert(0, REF_ROOT)
    try:
        util = importlib.import_module("models.pointnet_util")
        net = importlib.import_module("models.pointnet_extrusion")
        losses = importlib.import_module("losses")
        data_utils = importlib.import_module("data_utils")
    finally:
        # keep the reference modules reachable only through the returned namespace
        ref_mods = {}
        for k in list(saved_mods):
            if k in sys.modules:
                ref_mods[k] = sys.modules.pop(k)
            if saved_mods[k] is not None:
                sys.modules[k] = saved_mods[k]
        sys.path[:] = saved_path
    _loaded.update(util=util, net=net, losses=losses, data_utils=data_utils)
    return types.SimpleNamespace(**_loaded)


_loaded_igr = {}


def load_igr():
    """The reference's implicit sketch network: namespace with .network (IGR/network.py) and .sampler (IGR/sampler.py).
    Both resolve `general` through sys.path (train_Point2Cyl.py:17-21 appends IGR/); general.py imports trimesh at
    module scope only for mesh I/O helpers, so the stub suffices."""
    if _loaded_igr:
        return types.SimpleNamespace(**_loaded_igr)
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REF_ROOT}")
    _install_stubs()
    saved_path = list(sys.path)
    names = ("general", "network", "sampler")
    saved_mods = {k: sys.modules.get(k) for k in names}
    for k in names:
        sys.modules.pop(k, None)
    sys.path.insert(0, os.path.join(REF_ROOT, "IGR"))
    try:
        network = importlib.import_module("network")
        sampler = importlib.import_module("sampler")
    finally:
        for k in names:
            sys.modules.pop(k, None)
            if saved_mods[k] is not None:
                sys.modules[k] = saved_mods[k]
        sys.path[:] = saved_path
    _loaded_igr.update(network=network, sampler=sampler)
    return types.SimpleNamespace(**_loaded_igr)
