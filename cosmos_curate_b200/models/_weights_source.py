"""Where model weights come from (host side).

Order: an explicit directory -> the reference's weight cache (model_utils.get_local_dir_for_weights_name, the
lookup clip.py:39 / aesthetics.py:139 use) -> `CURATE_B200_WEIGHTS_DIR/<model id>` -> seeded synthetic weights,
ONLY when explicitly requested (`seed=` argument or CURATE_B200_SYNTHETIC_WEIGHTS=<seed>): a missing checkpoint is
an error, never a silent substitute.
"""

from __future__ import annotations

import os
from pathlib import Path


def resolve_dir(model_id: str, explicit: str | Path | None) -> Path | None:
    if explicit is not None:
        p = Path(explicit)
        if not p.is_dir():
            msg = f"weights directory {p} does not exist"
            raise FileNotFoundError(msg)
        return p
    try:  # full cosmos-curate environment
        from cosmos_curate.core.utils.model import model_utils

        p = Path(model_utils.get_local_dir_for_weights_name(model_id))
        if p.is_dir():
            return p
    except Exception:  # noqa: BLE001
        pass
    root = os.environ.get("CURATE_B200_WEIGHTS_DIR")
    if root:
        p = Path(root) / model_id
        if p.is_dir():
            return p
    return None


def synthetic_seed(explicit_seed: int | None) -> int | None:
    if explicit_seed is not None:
        return int(explicit_seed)
    env = os.environ.get("CURATE_B200_SYNTHETIC_WEIGHTS")
    return int(env) if env not in (None, "") else None
