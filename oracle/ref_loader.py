"""Stub loader that executes the reference's OWN hot-path files from /root/reference.  TEST INFRA.

In the authoring container the files are executed where they lie under ``/root/reference``; on the GPU box (no
``/root/reference``) from the staged copy ``oracle/_ref`` that ``oracle/build_ref.py`` produces (a git-ignored build
output).  Used by ``oracle/gen_golden.py``, ``oracle/ref_arm.py`` (the CPU arm of ``bench.py``) and the tests.

The reference cannot be imported as a package here (``omegaconf``, ``kornia``, ``lightning`` ...
are not installed, no network).  Its hot-path files load unmodified once
  * bare package modules (no ``__init__`` execution) are registered for ``lightning_pose`` and its
    sub-packages, so ``import lightning_pose.data.heatmaps`` resolves to the real file;
  * ``kornia`` is replaced by the six restated functions in ``oracle/lp_oracle.py``;
  * ``omegaconf`` / ``lightning.pytorch`` / three data-layer modules are placeholder names.
The sources are executed as they are (no edits); see oracle/build_ref.py for the staged copy.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")  # oracle/build_ref.py (GPU box)
REF_ROOT = os.environ.get("LP_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REF_ROOT, "lightning_pose")) and os.path.isdir(os.path.join(_STAGED, "lightning_pose")):
    REF_ROOT = _STAGED


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "lightning_pose"))


def _bare_package(name: str, path: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__path__ = [path]  # type: ignore[attr-defined]
    mod.__package__ = name
    sys.modules[name] = mod
    return mod


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def install() -> None:
    """Register stubs + bare packages (idempotent)."""
    if "lightning_pose" in sys.modules and getattr(sys.modules["lightning_pose"], "_lpb_stub", False):
        return
    import torch
    from oracle import lp_oracle as O

    # ---- kornia: the six functions on the path, restated in lp_oracle ---------------------
    def _filter2d(x, kernel, border_type="constant"):
        assert border_type == "constant" and tuple(kernel.shape) == (1, 5, 5)
        return O.pyramid_blur(x)

    def _pyr_kernel():
        k = torch.tensor([1.0, 4.0, 6.0, 4.0, 1.0])
        return (torch.outer(k, k) / 256.0)[None]

    def _kl(pred, target, reduction="none"):
        assert reduction == "none"
        return O.kl_div_2d(pred, target)

    def _js(pred, target, reduction="none"):
        assert reduction == "none"
        return O.js_div_2d(pred, target)

    _stub("kornia")
    _stub("kornia.filters", filter2d=_filter2d)
    _stub("kornia.geometry")
    _stub(
        "kornia.geometry.subpix",
        spatial_softmax2d=lambda x, temperature=torch.tensor(1.0): O.spatial_softmax2d(x, temperature),
        spatial_expectation2d=lambda p, normalized_coordinates=True: (
            O.spatial_expectation2d(p) if not normalized_coordinates else (_ for _ in ()).throw(NotImplementedError())
        ),
    )
    _stub("kornia.geometry.transform")
    _stub("kornia.geometry.transform.pyramid", _get_pyramid_gaussian_kernel=_pyr_kernel)
    _stub("kornia.losses", kl_div_loss_2d=_kl, js_div_loss_2d=_js)

    # ---- config / framework placeholders --------------------------------------------------
    class ListConfig(list):
        pass

    class DictConfig(dict):
        pass

    class OmegaConf:  # noqa: D401 - placeholder
        @staticmethod
        def to_object(x):
            return x

        @staticmethod
        def register_new_resolver(*a, **k):
            return None

    _stub("omegaconf", ListConfig=ListConfig, DictConfig=DictConfig, OmegaConf=OmegaConf)
    pl = _stub("lightning.pytorch", LightningModule=torch.nn.Module)
    _stub("lightning", pytorch=pl)

    # ---- bare packages so the real files are found without running package __init__ --------
    base = os.path.join(REF_ROOT, "lightning_pose")
    root = _bare_package("lightning_pose", base)
    root._lpb_stub = True  # type: ignore[attr-defined]
    for sub in ("data", "models", "models.heads", "models.backbones", "losses", "utils"):
        _bare_package("lightning_pose." + sub, os.path.join(base, *sub.split(".")))

    class _Placeholder:
        pass

    _stub("lightning_pose.data.datamodules", BaseDataModule=_Placeholder, UnlabeledDataModule=_Placeholder)
    _stub("lightning_pose.data.datasets", MultiviewHeatmapDataset=_Placeholder)
    _stub("lightning_pose.data.extractor", DataExtractor=_Placeholder)

    fac = importlib.import_module("lightning_pose.models.backbones.factory")
    sys.modules["lightning_pose.models.backbones"].BACKBONE_STRIDES = fac.BACKBONE_STRIDES  # type: ignore
    # heads/__init__ re-exports HeatmapHead (heatmap_mhcrnn.py imports it from the package)
    hm = importlib.import_module("lightning_pose.models.heads.heatmap")
    sys.modules["lightning_pose.models.heads"].HeatmapHead = hm.HeatmapHead  # type: ignore


def load(name: str) -> types.ModuleType:
    """Import a reference module (e.g. ``lightning_pose.models.heads.heatmap``) unmodified."""
    install()
    return importlib.import_module(name)
