#!/usr/bin/env python
"""Count the SASS instruction mix of the hottest loop (the one containing LDG.128) of a kernel.

usage: tools/sass_loop.py <lib.so> <mangled-substring>
No GPU needed: cuobjdump -sass on the cross-compiled library.
"""
import collections
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
MARK = sys.argv[3] if len(sys.argv) > 3 else "LDG"  # LDG: direct-load loop, LDS: cp.async ring loop
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    # backward branches = loops
    loops = []
    for i, (a, s) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", s)
        if m:
            t = int(m.group(1), 16)
            if t < a and t in addr:
                loops.append((addr[t], i))
    best = None
    for lo, hi in loops:
        n128 = sum(1 for _, s in ins[lo:hi + 1] if ("LDG" in s and ".128" in s and "LDGSTS" not in s) == (MARK == "LDG")
                   and (MARK != "LDS" or ("LDS" in s and ".128" in s)))
        if n128 and (best is None or hi - lo < best[1] - best[0]):
            best = (lo, hi, n128)  # innermost loop that streams pixels
    print(f"== {name}: {len(ins)} instrs total")
    if not best:
        print("   no LDG.128 loop found")
        continue
    lo, hi, n128 = best
    body = [s for _, s in ins[lo:hi + 1]]
    cnt = collections.Counter()
    for s in body:
        s = re.sub(r"^@!?U?P\d+\s+", "", s)
        op = s.split()[0].split(".")[0]
        cnt[op] += 1
    px = n128 * 4 / 5.0  # 5 planes per 4 pixels
    print(f"   loop [{ins[lo][0]:#x},{ins[hi][0]:#x}] {len(body)} instrs, {n128} {MARK}.128 -> {px:.0f} pixels/iter, {len(body)/px:.1f} instr/pixel")
    fma = sum(cnt[k] for k in ("FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2"))
    alu = sum(cnt[k] for k in ("FMNMX", "FSEL", "FSETP", "IADD3", "IADD", "LOP3", "SHF", "IMAD", "MOV", "ISETP", "SEL", "LEA", "FMNMX3", "PRMT", "I2FP", "VIADD", "UMOV", "HFMA2"))
    print(f"   fp32-pipe {fma} ({fma/px:.1f}/px)  mufu {cnt['MUFU']} ({cnt['MUFU']/px:.1f}/px)  other-alu {alu} ({alu/px:.1f}/px)")
    # FP32-pipe cost model from tools/ubench_fp32x2.cu (1 / measured warp-instr per clk per SM)
    cost = 0.65 * cnt["FFMA2"] + 0.306 * (cnt["FMUL2"] + cnt["FADD2"]) + 0.37 * (cnt["FFMA"] + cnt["FMUL"] + cnt["FADD"])
    print(f"   fp32-pipe cost model: {cost/px:.1f} clk/32px  (HBM roofline = 28.3 clk/32px at 1965 MHz)")
    print("   " + ", ".join(f"{k}:{v}" for k, v in cnt.most_common()))
