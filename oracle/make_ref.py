#!/usr/bin/env python
"""TEST INFRASTRUCTURE — not product code.

Vendors the two reference modules of the hot path (models/cxrbert_origin.py, models/image.py) from /root/reference into
oracle/_ref/ — a BUILT artefact: git-ignored, never committed, but shipped to the GPU box with the working tree — so that
`bench.py --impl reference` and the `cpu_baseline` leg can time the reference's OWN modules on the GPU box's host cores
(`cpu_baseline.kind = "reference"`) instead of the oracle port.  The files are copied byte for byte; they
run under the compatibility shims of oracle/ref_shim.py (old transformers paths, offline BertConfig, random-init ResNet).
__graft_entry__.build() calls this whenever /root/reference is present; nothing in the product imports oracle/_ref."""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MEDVILL_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("models/cxrbert_origin.py", "models/image.py")


def main():
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        print("make_ref: %s not present; oracle/_ref left as is" % SRC)
        return 0
    man = []
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        man.append("%s  %s" % (hashlib.sha256(open(dst, "rb").read()).hexdigest(), rel))
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(man) + "\n")
    print("make_ref: vendored %d reference modules into %s" % (len(FILES), DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
