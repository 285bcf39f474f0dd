"""Import the reference's OWN hot-path modules from /root/reference (build container only).

Used by oracle/make_golden.py to produce tests/golden/* and by tests that pin the oracle when the
checkout is present.  /root/reference does not exist on the GPU box: nothing that runs there may
call into this module (tests skip when `available()` is False).

Two stubs are injected, as established in SURVEY.md 8c / V2:
  * ``av``  (PyAV is not installed; only needed at import time by decoder_utils), and
  * ``cosmos_curate.core.utils.model.model_utils`` (weight-directory lookup; one function).

Test infrastructure only (see oracle/__init__.py).
"""

from __future__ import annotations

import os
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path(os.environ.get("CURATE_REFERENCE_ROOT", "/root/reference"))
_weights_dirs: dict[str, Path] = {}


def available() -> bool:
    return (REFERENCE_ROOT / "cosmos_curate" / "models" / "clip.py").exists()


def register_weights_dir(model_id: str, path: Path) -> None:
    _weights_dirs[model_id] = Path(path)


def _install_stubs() -> None:
    os.environ.setdefault("CONDA_DEFAULT_ENV", "unified")
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    if "av" not in sys.modules:
        try:
            import av  # noqa: F401
        except ImportError:
            av_stub = types.ModuleType("av")
            av_stub.VideoFrame = object  # names referenced in annotations only
            container = types.ModuleType("av.container")
            container.InputContainer = object
            av_stub.container = container
            sys.modules["av"] = av_stub
            sys.modules["av.container"] = container
    name = "cosmos_curate.core.utils.model.model_utils"
    if name not in sys.modules:
        stub = types.ModuleType(name)

        def get_local_dir_for_weights_name(weights_name: str) -> Path:
            return _weights_dirs[weights_name]

        stub.get_local_dir_for_weights_name = get_local_dir_for_weights_name
        sys.modules[name] = stub


def decoder_utils():
    _install_stubs()
    import importlib

    return importlib.import_module("cosmos_curate.pipelines.video.utils.decoder_utils")


def clip_module():
    _install_stubs()
    import importlib

    return importlib.import_module("cosmos_curate.models.clip")


def aesthetics_module():
    _install_stubs()
    import importlib

    return importlib.import_module("cosmos_curate.models.aesthetics")
