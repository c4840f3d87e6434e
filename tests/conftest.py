import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """GPU sessions: write the max relative errors seen by every assert_close (softened and plain) to gpurun_out/."""
    try:
        import json
        from tests import helpers
        import torch
        if helpers.ERROR_LOG and torch.cuda.is_available():
            out = ROOT / "gpurun_out"
            out.mkdir(exist_ok=True)
            rows = {k: dict(softened=v[0], plain_masked=v[1], calls=v[2]) for k, v in sorted(helpers.ERROR_LOG.items())}
            (out / "parity_errors.json").write_text(json.dumps(rows, indent=1))
    except Exception:                                            # pragma: no cover
        pass
