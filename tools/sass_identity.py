"""Compare the SASS of every kernel in two object files (whitespace and line-info comments ignored): used to show that
adding an experimental kernel to a translation unit left the shipped kernels' machine code untouched.
usage: python tools/sass_identity.py old.o new.o"""
import re
import subprocess
import sys


def funcs(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for ln in txt.split("\n"):
        if ln.strip().startswith("//##"):
            continue
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            out[cur] = []
        elif cur:
            out[cur].append(" ".join(ln.split()))
    return out


if __name__ == "__main__":
    a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
    bad = 0
    for k in a:
        same = a[k] == b.get(k)
        bad += not same
        print("%-70s %s" % (k[:70], "identical" if same else "DIFFERENT"))
    print("only in new:", [k for k in b if k not in a])
    sys.exit(1 if bad else 0)
