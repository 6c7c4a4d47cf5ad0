#!/usr/bin/env python
"""development aid: md5 of the SASS of every kernel in a built libstepsb200.so (addresses included: identical code gives an
identical listing).  Used to show that a change which cannot be re-run on a GPU leaves the kernels verified there untouched.
usage: sass_fingerprint.py lib.so > fingerprints.txt ;  sass_fingerprint.py old.txt new.txt  (compare)"""
import hashlib
import re
import subprocess
import sys


def fingerprints(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
    fp, name, buf = {}, None, []
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            if name:
                fp[name] = hashlib.md5("\n".join(buf).encode()).hexdigest()
            name, buf = m.group(1), []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln):
            buf.append(ln.strip())
    if name:
        fp[name] = hashlib.md5("\n".join(buf).encode()).hexdigest()
    return fp


def load(path):
    return dict(ln.split() for ln in open(path) if ln.strip())


if __name__ == "__main__":
    if len(sys.argv) == 2:
        for k, v in sorted(fingerprints(sys.argv[1]).items()):
            print(k, v)
    else:
        a, b = load(sys.argv[1]), load(sys.argv[2])
        same = [k for k in a if k in b and a[k] == b[k]]
        diff = [k for k in a if k in b and a[k] != b[k]]
        gone = [k for k in a if k not in b]
        new = [k for k in b if k not in a]
        print(f"{len(same)} kernels identical, {len(diff)} changed, {len(gone)} removed, {len(new)} new")
        for k in diff:
            print("CHANGED", k)
        for k in gone:
            print("REMOVED", k)
        for k in new:
            print("NEW", k)
