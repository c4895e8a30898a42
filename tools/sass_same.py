"""Compare the SASS of two object files kernel by kernel, ignoring trailing bool template arguments added to the names:
    python tools/sass_same.py old.o new.o
Used to show that adding a template-specialised variant (split-K, dropout) left the default instantiations' instruction streams
untouched when no GPU is at hand to re-measure them."""
import re
import subprocess
import sys


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    d, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
        elif cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            d[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip())
    return d


def main(old, new):
    a, b = kernels(old), kernels(new)
    bad = 0
    for k, v in b.items():
        def stem(n, strip):
            n = n.split("EEv")[0]                                 # template name + arguments, without the parameter types
            return re.sub(r"Lb0E$", "", n) if strip else n
        hit = next((o for strip in (False, True) for o in a if stem(o, False) == stem(k, strip)), None)
        if hit is None:
            print("new      %6d  %s" % (len(v), k[:110]))
        elif a[hit] == v:
            print("same     %6d  %s" % (len(v), k[:110]))
        else:
            bad += 1
            print("DIFFERS  %6d -> %6d  %s" % (len(a[hit]), len(v), k[:110]))
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(sys.argv[1], sys.argv[2]) else 0)
