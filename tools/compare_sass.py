"""Builds the library of another commit in a scratch directory and compares the SASS of every kernel with the current
build, kernel by kernel (instruction text without addresses / encodings).  Used at the end of round 1 to show that the
kernels of the default path are byte-for-byte the ones last verified on hardware (commit 3ee5c68) after the
experimental variants were added next to them.  Needs no GPU.   python tools/compare_sass.py <commit>"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from densesurfelmapping_b200 import build as B  # noqa: E402


def sass(so):
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    out, name = collections.defaultdict(list), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", line)
        if m and name:
            out[name].append(m.group(1))
    return out


def main():
    commit = sys.argv[1]
    with tempfile.TemporaryDirectory() as tmp:
        tar = subprocess.run(["git", "-C", ROOT, "archive", commit, "densesurfelmapping_b200/csrc", "include"], capture_output=True, check=True).stdout
        subprocess.run(["tar", "-x", "-C", tmp], input=tar, check=True)
        src = os.path.join(tmp, "densesurfelmapping_b200", "csrc")
        files = [os.path.join(src, f) for f in sorted(os.listdir(src)) if f.endswith((".cu", ".cpp"))]
        old_so = os.path.join(tmp, "old.so")
        subprocess.run([B._nvcc()] + B.NVCC_FLAGS + ["-o", old_so] + files, check=True, capture_output=True)
        old, new = sass(old_so), sass(B.build())
    for k in sorted(old):
        print(("SAME " if old[k] == new.get(k) else "DIFF ") + k, len(old[k]), len(new.get(k, [])))
    print("only in the current build:", sorted(set(new) - set(old)))


if __name__ == "__main__":
    main()
