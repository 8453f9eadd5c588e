"""Import the reference's own modules verbatim in the authoring container (TEST INFRASTRUCTURE ONLY).

``/root/reference`` exists only in the authoring container, never on the GPU box, so nothing in ``tests -m gpu``,
``smoke()`` or ``bench.py`` uses this file; it is used by ``oracle/make_golden.py`` to pin the oracle and to
generate ``tests/golden/*``.

The reference imports packages that are absent here (monai, pytorch_lightning, itk, matplotlib, its sibling
``transforms``).  They are replaced by permissive stubs; the only *arithmetic* stand-ins are
``monai.networks.nets.UNet`` (-> ``oracle.monai_unet.UNet``) and a functional
``RandSpatialCropSamplesd``/``Compose`` (-> slicing with injected origins, see ``oracle/gan.py``).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

from . import monai_unet

REFERENCE_ROOT = "/root/reference"
# oracle/build_ref.py copies the two reference files of the path here (git-ignored; travels to the GPU box with the
# snapshot) so that `bench.py --impl reference` can time the reference's OWN classes there
REF_COPY_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


class _Anything:
    """Attribute/call sink for symbols the hot path never executes."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _Hparams(dict):
    __getattr__ = dict.__getitem__


class LightningModule(nn.Module):
    """Minimal stand-in: save_hyperparameters / hparams / log."""

    def __init__(self, *a, **k):
        super().__init__()
        self.hparams = _Hparams()
        self.logged = {}

    def save_hyperparameters(self, *names):
        import inspect
        frame = inspect.currentframe().f_back
        for n in names:
            self.hparams[n] = frame.f_locals[n]

    def log(self, name, value, **kw):
        self.logged[name] = value.detach().clone() if torch.is_tensor(value) else value


class RandSpatialCropSamplesd:
    """Functional stand-in; origins are injected through ``RandSpatialCropSamplesd.origins`` (B, S, dims)."""
    origins = None

    def __init__(self, keys, roi_size, num_samples, random_size=False):
        self.keys, self.roi, self.num_samples = keys, roi_size, num_samples

    def __call__(self, data, vol_idx):
        out = []
        for s in range(self.num_samples):
            o = type(self).origins[vol_idx][s]
            sl = (slice(None),) + tuple(slice(int(x), int(x) + r) for x, r in zip(o, self.roi))
            out.append({k: data[k][sl] for k in self.keys})
        return out


class Compose:
    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, data_list):
        res = []
        for i, d in enumerate(data_list):
            for t in self.transforms:
                d = t(d, i)
            res.append(d)
        return res


def _install_stubs():
    names = ["monai", "monai.apps", "monai.config", "monai.data", "monai.inferers", "monai.losses",
             "monai.metrics", "monai.networks", "monai.networks.layers", "monai.networks.nets",
             "monai.transforms", "monai.utils", "monai.visualize", "monai.visualize.img2tensorboard",
             "pytorch_lightning", "pytorch_lightning.loggers", "pytorch_lightning.callbacks",
             "pytorch_lightning.callbacks.model_checkpoint", "itk", "matplotlib", "matplotlib.pyplot",
             "transforms", "torchvision", "torchvision.transforms"]
    saved = {n: sys.modules.get(n) for n in names}
    for n in names:
        sys.modules[n] = _StubModule(n)
    for n in names:  # parent.child attribute links
        if "." in n:
            parent, child = n.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[n])
    sys.modules["monai.networks.nets"].UNet = monai_unet.UNet
    norm = types.SimpleNamespace(BATCH="batch", INSTANCE="instance")
    sys.modules["monai.networks.layers"].Norm = norm
    sys.modules["monai.transforms"].RandSpatialCropSamplesd = RandSpatialCropSamplesd
    sys.modules["monai.transforms"].Compose = Compose
    pl = sys.modules["pytorch_lightning"]
    pl.LightningModule = LightningModule
    pl.LightningDataModule = object
    return saved


def _restore(saved):
    for n, m in saved.items():
        if m is None:
            sys.modules.pop(n, None)
        else:
            sys.modules[n] = m


def load_reference_module(relpath, name, root=None):
    """Execute a reference file verbatim (e.g. 'code/GAN/GAN_final.py') and return it as a module.  ``root``: the
    reference tree (default /root/reference, else the oracle/_ref copy)."""
    if root is None:
        root = REFERENCE_ROOT if os.path.exists(os.path.join(REFERENCE_ROOT, relpath)) else REF_COPY_ROOT
    path = os.path.join(root, relpath)
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    saved = _install_stubs()
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        _restore(saved)
    return mod


def available(root=None):
    roots = (root,) if root else (REFERENCE_ROOT, REF_COPY_ROOT)
    return any(os.path.exists(os.path.join(r, "code/GAN/GAN_final.py")) for r in roots)
