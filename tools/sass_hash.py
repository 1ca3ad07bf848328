#!/usr/bin/env python3
"""Per-function instruction count and hash of the SASS in an object file (cuobjdump -sass, addresses stripped):
   python tools/sass_hash.py kspace_neutrinos_b200/build/k1_powerspec.o
Used to show that adding an opt-in template instantiation leaves the kernels already measured on the B200 byte-identical
(compare the lines of the old and the new object)."""
import sys, re, hashlib, subprocess
def funcs(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, d = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); d[cur] = []
        elif cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
            d[cur].append(re.sub(r"/\*[0-9a-f]+\*/", "", line).strip())
    return {k: (len(v), hashlib.md5("\n".join(v).encode()).hexdigest()[:10]) for k, v in d.items()}
if __name__ == "__main__":
    f = funcs(sys.argv[1])
    names = subprocess.run(["c++filt"], input="\n".join(f), capture_output=True, text=True).stdout.splitlines()
    for (k, v), n in zip(f.items(), names):
        print(v[0], v[1], n[:140])
