"""Import shim for the UNMODIFIED reference package (`label_anything`).

BASELINE / TEST INFRASTRUCTURE ONLY -- never imported by labelanything_b200.  Two users:
  * bench.py --impl reference / the cpu_baseline leg: imports the reference installed (unmodified, `pip install
    --no-deps --target baseline/_ref`) under baseline/_ref/, which travels to the GPU box;
  * oracle/ref_import.py -> oracle/make_golden.py: imports it from /root/reference in the build container.
The reference eagerly imports packages its hot path never uses (ruamel.yaml, colorlog, matplotlib, pycocotools, timm,
dropblock, ...); those that are absent get inert stub modules.  `transformers` must be imported BEFORE the stubs are
installed.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import sys
import types

from pathlib import Path

INSTALLED_ROOT = str(Path(__file__).resolve().parent / "_ref")   # pip --target of the unmodified reference
_STUB_ROOTS = ("ruamel", "colorlog", "matplotlib", "pycocotools", "timm", "dropblock", "accelerate",
               "torchmetrics", "captum", "easydict", "optuna", "wandb", "cv2", "lvis", "nicegui", "streamlit")


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __mro_entries__(self, bases):
        return (object,)

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root in _STUB_ROOTS:
            try:
                # only stub what is genuinely missing
                for f in sys.meta_path:
                    if f is self:
                        continue
                    spec = f.find_spec(fullname, path, target) if hasattr(f, "find_spec") else None
                    if spec is not None:
                        return None
            except Exception:
                pass
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def import_reference(root: str = INSTALLED_ROOT):
    """Returns the reference's `label_anything.models` package, imported from `root`."""
    if not (Path(root) / "label_anything" / "models" / "lam.py").exists():
        raise ImportError(f"the reference package is not installed under {root}")
    global _installed
    import transformers  # noqa: F401  (must precede the stubs)
    import transformers.models.vit.modeling_vit  # noqa: F401

    if not _installed:
        sys.meta_path.append(_StubFinder())
        if root not in sys.path:
            sys.path.insert(0, root)
        _installed = True
    import label_anything.models as models

    return models
