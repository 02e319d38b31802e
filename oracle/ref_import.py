"""Import the UNMODIFIED reference from /root/reference in the build container (fixture generation only).

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py to generate the committed fixtures under tests/golden/.
/root/reference does not exist on the GPU box: nothing in tests -m gpu / smoke / bench imports this file.  The stub
machinery for the reference's unrelated imports lives in baseline/ref_shim.py (shared with bench.py's reference arm).
"""
from __future__ import annotations

import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from baseline import ref_shim  # noqa: E402

REFERENCE_ROOT = "/root/reference"


def import_reference():
    """Returns the reference's `label_anything.models` package."""
    return ref_shim.import_reference(REFERENCE_ROOT)
