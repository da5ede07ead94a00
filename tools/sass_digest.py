#!/usr/bin/env python3
"""Per-kernel digest of the SASS in gr_dvbt_b200/libdvbt_b200.so: `sha1  #instructions  demangled-ish name`.

Use: a session without GPU access can still prove that a source change left the GPU-verified kernels alone -
compare the digest of the new build with the committed digest of the last build whose `pytest -m gpu` run was green
(profiles/rNN_sass_digest_*.txt):

    python tools/sass_digest.py > /tmp/new.txt && python tools/sass_digest.py --diff profiles/r01_sass_digest_v37.txt /tmp/new.txt

Instruction text only (addresses and encodings stripped, which also removes nothing semantic: operands, predicates
and constant-bank offsets stay).  The anonymous-namespace hash in the mangled names is removed so that digests of
different builds of the same sources line up."""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gr_dvbt_b200", "libdvbt_b200.so")


def kernels(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"_GLOBAL__N__[0-9a-f]+_", "_GLOBAL__N__", m.group(1))
            out[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?;)", line)
        if cur and m:
            out[cur].append(m.group(1).strip())
    return out


def digest(lib):
    rows = []
    for name, ins in sorted(kernels(lib).items()):
        rows.append("%s %6d %s" % (hashlib.sha1("\n".join(ins).encode()).hexdigest()[:16], len(ins), name))
    return rows


def load(path):
    d = {}
    for line in open(path):
        if line.startswith("#") or not line.strip():
            continue
        h, n, name = line.split(None, 2)
        d[name.strip()] = (h, int(n))
    return d


def main():
    if len(sys.argv) >= 4 and sys.argv[1] == "--diff":
        a, b = load(sys.argv[2]), load(sys.argv[3])
        same = [k for k in a if k in b and a[k] == b[k]]
        changed = [k for k in a if k in b and a[k] != b[k]]
        removed = [k for k in a if k not in b]
        added = [k for k in b if k not in a]
        # a kernel whose name changed (a template parameter's type, say) but whose code did not
        renamed = []
        for k in list(removed):
            twin = [j for j in added if b[j] == a[k]]
            if twin:
                renamed.append((k, twin[0]))
                removed.remove(k)
                added.remove(twin[0])
        print("identical: %d   renamed, same code: %d   changed: %d   removed: %d   added: %d" % (
            len(same), len(renamed), len(changed), len(removed), len(added)))
        for k in changed:
            print("CHANGED", k, a[k], "->", b[k])
        for k, j in renamed:
            print("RENAMED", k, "->", j)
        for k in removed:
            print("REMOVED", k)
        for k in added:
            print("ADDED  ", k)
        return
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    print("# tools/sass_digest.py of %s" % os.path.relpath(lib, ROOT))
    print("\n".join(digest(lib)))


if __name__ == "__main__":
    main()
