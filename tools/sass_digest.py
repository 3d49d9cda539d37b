#!/usr/bin/env python
"""tools/sass_digest.py <lib.so> [<other.so>]: md5 of the SASS (opcodes + operands, no addresses / encodings) of
every kernel; with two libraries, list the kernels whose code differs -- proves that a refactoring left the
generated device code alone."""
import hashlib, re, subprocess, sys

def digests(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    d, name, h = {}, None, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name: d[name] = (h.hexdigest(), n)
            name, h, n = m.group(1), hashlib.md5(), 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", line)
        if m and name:
            h.update(m.group(1).encode()); n += 1
    if name: d[name] = (h.hexdigest(), n)
    return d

a = digests(sys.argv[1])
if len(sys.argv) == 2:
    for k, (v, n) in sorted(a.items()): print(v, n, k)
else:
    b = digests(sys.argv[2])
    diff = [k for k in sorted(set(a) | set(b)) if a.get(k) != b.get(k)]
    print(f"{len(a)} / {len(b)} kernels, {len(diff)} differ")
    for k in diff: print("  ", k, a.get(k, ("-", 0))[1], "->", b.get(k, ("-", 0))[1], "instructions")
