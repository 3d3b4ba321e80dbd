"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the *unmodified* CFUN reference from /root/reference so the oracle
restatement (oracle/cfun_oracle.py) can be pinned against it and so
oracle/gen_golden.py can emit golden vectors.  /root/reference exists only in
the build container, never on the GPU box, so nothing under tests -m gpu,
smoke() or bench.py may import this module.

Shims: (1) empty stand-ins for the absent third-party imports
`nibabel`, `skimage`, `skimage.transform` (reference utils.py:9-10,
heart_main.py:13); `skimage.transform.resize` is served by
scipy.ndimage.zoom(order=0/1, mode='grid-constant', grid_mode=True), the call skimage >= 0.19
delegates to ("parity unpinned" at that third-party boundary: scikit-image is
unpinned in reference README.md:16); (2) Tensor.cuda / Module.cuda become the
identity so the reference's unconditional .cuda() calls run on CPU.
"""
import os
import sys
import types
import importlib
import contextlib

REF_ROOT = os.environ.get("CFUN_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "model.py"))


def _skimage_resize(image, output_shape, order=1, mode="constant", cval=0, clip=True,
                    preserve_range=True, anti_aliasing=False, anti_aliasing_sigma=None):
    import numpy as np
    import scipy.ndimage as ndi
    image = np.asarray(image)
    out_shape = tuple(int(s) for s in output_shape)
    zoom = [o / i for o, i in zip(out_shape, image.shape)]
    # skimage>=0.19: resize -> ndi.zoom(..., mode=_to_ndimage_mode('constant') == 'grid-constant', grid_mode=True)
    res = ndi.zoom(image.astype(np.float64), zoom, order=order, mode="grid-constant", cval=cval, grid_mode=True)
    assert res.shape == out_shape, (res.shape, out_shape)
    return res


_loaded = {}
_loaded_by_root = {}


def load(subdir=None):
    """Returns dict of reference modules: model, utils, backbone, mask_branch, config.  subdir="LiTS_2017" loads the liver
    variant (same module names, its own directory) instead of the heart tree."""
    if subdir:
        return _load_from(os.path.join(REF_ROOT, subdir), "cfunref_%s_" % subdir.lower())
    if _loaded:
        return _loaded
    _loaded.update(_load_from(REF_ROOT, "cfunref_"))
    return _loaded


def _load_from(ref_root, tag):
    if ref_root in _loaded_by_root:
        return _loaded_by_root[ref_root]
    _loaded = {}
    REF_ROOT = ref_root
    if not os.path.isfile(os.path.join(REF_ROOT, "model.py")):
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import torch
    import torch.nn as nn
    sys.dont_write_bytecode = True
    for name in ("nibabel", "skimage", "skimage.transform"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sk = sys.modules["skimage"]
    sk.__version__ = "0.19.0"
    sk.transform = sys.modules["skimage.transform"]
    sk.transform.resize = _skimage_resize
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    # the repo root holds same-named drop-in shims (model.py, utils.py ...); make
    # sure the names resolve to the reference tree for the duration of the import.
    saved = {k: sys.modules.pop(k) for k in ("model", "utils", "backbone", "mask_branch", "config")
             if k in sys.modules}
    sys.path.insert(0, REF_ROOT)
    try:
        mods = {k: importlib.import_module(k) for k in ("utils", "backbone", "mask_branch", "model", "config")}
    finally:
        sys.path.remove(REF_ROOT)
        for k in ("model", "utils", "backbone", "mask_branch", "config"):
            m = sys.modules.pop(k, None)
            if m is not None:
                sys.modules[tag + k] = m
        sys.modules.update(saved)
    _loaded.update(mods)
    _loaded_by_root[ref_root] = _loaded
    return _loaded


@contextlib.contextmanager
def quiet():
    """The reference prints from inside its layers (model.py:447,1373); mute it."""
    with open(os.devnull, "w") as dn, contextlib.redirect_stdout(dn):
        yield
