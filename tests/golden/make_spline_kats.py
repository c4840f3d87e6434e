"""Writes tests/golden/spline_kats.json: known-answer vectors for the degree-1 open B-spline basis
(torch_spline_conv's published algorithm), hand-derived in SURVEY.md section 8(c).  The reference ships
no golden vectors (no tests at all) and torch_spline_conv is not installable offline,
so these KATs are the only fixed points for the third-party arithmetic ("parity unpinned")."""
import json
from pathlib import Path

kats = {
    "kernel_size": 5,
    "one_d": {
        "pseudo": [0.0, 0.0625, 0.25, 0.75, 0.9375, 1.0],
        "basis": [[1, 0], [0.75, 0.25], [1, 0], [1, 0], [0.25, 0.75], [1, 0]],
        "index": [[0, 1], [0, 1], [1, 2], [3, 4], [3, 4], [4, 0]],
    },
    "two_d": {
        "pseudo": [[0.125, 0.5], [0.5, 0.5], [0.75, 0.125]],
        "basis": [[0.5, 0.5, 0, 0], [1, 0, 0, 0], [0.5, 0, 0.5, 0]],
        "index": [[10, 11, 15, 16], [12, 13, 17, 18], [3, 4, 8, 9]],
    },
}
Path(__file__).with_name("spline_kats.json").write_text(json.dumps(kats, indent=1))
