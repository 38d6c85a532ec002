"""Import shim for the *real* reference package (test infrastructure only).

The reference (`/root/reference/contrast_gan_3D`) is pure Python but pulls a few
optional third-party packages at import time that are not installed in this image
(`batchgenerators`, `patchly`, `matplotlib`, `seaborn`, `SimpleITK`, `torchio`,
`wandb`...).  None of them touches the arithmetic of the hot path (reference
`alias.py:7-13`, `trainer/logger/WandbLogger.py:6-8`, `utils/io_utils.py:5-6`,
`eval/CCTAContrastCorrector.py:7-8`), so we register empty stand-ins in
`sys.modules` and import the reference modules unmodified.

`tests/golden/make_golden.py`, the reference-vs-oracle CPU tests and `bench.py`'s reference
arm / `cpu_baseline` leg use this.  `/root/reference` does not exist on the GPU box: there the
shim resolves to `oracle/_ref/` (a verbatim copy made by `oracle/make_ref.py`).
"""
from __future__ import annotations

import importlib
import importlib.machinery
import sys
import types
from pathlib import Path

# The reference tree itself in the authoring container; on the GPU box (no /root/reference) the verbatim copy made by
# oracle/make_ref.py (git-ignored, shipped with the snapshot).
_LIVE = Path("/root/reference")
_COPY = Path(__file__).resolve().parent / "_ref"
REFERENCE_ROOT = _LIVE if (_LIVE / "contrast_gan_3D").is_dir() else _COPY

_STUBS = {
    "batchgenerators": [],
    "batchgenerators.dataloading": [],
    "batchgenerators.dataloading.multi_threaded_augmenter": ["MultiThreadedAugmenter"],
    "batchgenerators.dataloading.nondet_multi_threaded_augmenter": ["NonDetMultiThreadedAugmenter"],
    "batchgenerators.dataloading.single_threaded_augmenter": ["SingleThreadedAugmenter"],
    "batchgenerators.dataloading.data_loader": ["DataLoader"],
    "batchgenerators.transforms": [],
    "batchgenerators.transforms.abstract_transforms": ["Compose"],
    "batchgenerators.transforms.spatial_transforms": ["SpatialTransform_2"],
    "batchgenerators.transforms.utility_transforms": ["NumpyToTensor"],
    "batchgenerators.augmentations": [],
    "batchgenerators.augmentations.crop_and_pad_augmentations": ["crop"],
    "batchgenerators.augmentations.utils": ["pad_nd_image"],
    "batchgenerators.utilities": [],
    "batchgenerators.utilities.file_and_folder_operations": ["load_pickle", "write_pickle"],
    "patchly": [],
    "patchly.aggregator": ["Aggregator"],
    "patchly.sampler": ["GridSampler"],
    "matplotlib": ["colormaps"],
    "matplotlib.pyplot": [],
    "matplotlib.figure": ["Figure"],
    "matplotlib.axes": ["Axes"],
    "matplotlib.colors": ["Normalize"],
    "matplotlib.cm": [],
    "matplotlib.patches": ["Patch"],
    "seaborn": [],
    "SimpleITK": [],
    "torchio": [],
    "wandb": [],
    "torchvision": [],
    "torchvision.utils": ["make_grid"],
    "h5py": [],
    "openpyxl": [],
}


class _AnyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _AnyMeta(name, (_Anything,), {})


class _Anything(metaclass=_AnyMeta):
    """Placeholder class: constructible, callable, attribute access returns itself."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


def _make_stub(name: str, attrs):
    mod = types.ModuleType(name)
    mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    mod.__path__ = []  # behave like a package so sub-imports resolve
    for a in attrs:
        setattr(mod, a, _AnyMeta(a, (_Anything,), {}))

    def _getattr(attr, _n=name):
        if attr.startswith("__"):
            raise AttributeError(attr)
        sub = sys.modules.get(f"{_n}.{attr}")
        if sub is not None:
            return sub
        return _AnyMeta(attr, (_Anything,), {})

    mod.__getattr__ = _getattr
    return mod


def available() -> bool:
    return (REFERENCE_ROOT / "contrast_gan_3D" / "model" / "generator.py").is_file()


def install() -> None:
    """Register stand-ins for missing optional deps and put the reference on sys.path."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name, attrs in _STUBS.items():
        try:
            if name not in sys.modules:
                importlib.import_module(name)
        except Exception:
            sys.modules[name] = _make_stub(name, attrs)
    root = str(REFERENCE_ROOT)
    if root not in sys.path:
        sys.path.insert(0, root)


def load():
    """Return a namespace with the reference symbols on the hot path."""
    install()
    ns = types.SimpleNamespace()
    ns.blocks = importlib.import_module("contrast_gan_3D.model.blocks")
    ns.generator = importlib.import_module("contrast_gan_3D.model.generator")
    ns.discriminator = importlib.import_module("contrast_gan_3D.model.discriminator")
    ns.loss = importlib.import_module("contrast_gan_3D.model.loss")
    ns.model_utils = importlib.import_module("contrast_gan_3D.model.utils")
    ns.scaler = importlib.import_module("contrast_gan_3D.data.Scaler")
    ns.alias = importlib.import_module("contrast_gan_3D.alias")
    ns.constants = importlib.import_module("contrast_gan_3D.constants")
    return ns


def load_trainer():
    install()
    return importlib.import_module("contrast_gan_3D.trainer.Trainer")
