#!/usr/bin/env python
"""Build kernel variants for A/B timing on the GPU box.

    python tools/ab_build.py [--only conv_l1.cu] NAME "-DFOO=1 -DBAR=2" [NAME2 "..."]   ->  dagr_b200/build/variants/NAME.so

With --only, just that source is recompiled per variant (all variants in parallel) and linked with the objects of the
regular in-tree build (python -m dagr_b200.build).

and on the box:  DAGR_B200_LIB=dagr_b200/build/variants/NAME.so python bench.py ...
"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from dagr_b200 import build as B  # noqa: E402


def main(argv):
    out = B.PKG / "build" / "variants"
    out.mkdir(parents=True, exist_ok=True)
    only = None
    if argv and argv[0] == "--only":
        only, argv = argv[1], argv[2:]
    jobs = []
    for name, defs in zip(argv[0::2], argv[1::2]):
        objs, procs = [], []
        for s in B.SOURCES:
            if only and s != only:
                objs.append(str(B.PKG / "build" / s.replace(".cu", ".o")))
                continue
            o = out / f"{name}_{s.replace('.cu', '.o')}"
            procs.append(subprocess.Popen([B._nvcc(), "-c", str(B.CSRC / s), "-o", str(o)] + B.NVCC_FLAGS + defs.split(),
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
            objs.append(str(o))
        jobs.append((name, objs, procs))
    for name, objs, procs in jobs:
        logs = [p.communicate()[0] for p in procs]
        if any(p.returncode for p in procs):
            raise SystemExit("\n".join(logs))
        (out / f"{name}.ptxas.log").write_text("\n".join(logs))
        subprocess.check_call([B._nvcc(), "-shared", "-o", str(out / f"{name}.so")] + objs +
                              ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
        print(out / f"{name}.so")


if __name__ == "__main__":
    main(sys.argv[1:])
