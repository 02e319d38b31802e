"""Import shim for the UNMODIFIED reference (/root/reference) in the build container.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py to generate the committed fixtures under
tests/golden/.  The reference eagerly imports six packages its hot path never uses (ruamel.yaml,
colorlog, matplotlib, pycocotools, timm, dropblock); they are absent here, so we register inert stub
modules for them.  `transformers` must be imported BEFORE the stubs are installed.
/root/reference does not exist on the GPU box: nothing in tests -m gpu / smoke / bench imports this file.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import sys
import types

REFERENCE_ROOT = "/root/reference"
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


def import_reference():
    """Returns the reference's `label_anything.models` package."""
    global _installed
    import transformers  # noqa: F401  (must precede the stubs)
    import transformers.models.vit.modeling_vit  # noqa: F401

    if not _installed:
        sys.meta_path.append(_StubFinder())
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        _installed = True
    import label_anything.models as models

    return models
