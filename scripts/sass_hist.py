"""SASS instruction histogram of the loop bodies of the kernels whose name matches a pattern.
usage: sass_hist.py <object or cubin> <name regex>   (cuobjdump -sass; loops = backward branches)"""
import collections
import re
import subprocess
import sys

txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
pat = re.compile(sys.argv[2])
for f in re.split(r"\n\s+Function : ", txt)[1:]:
    name = f.split("\n")[0]
    if not pat.search(name):
        continue
    ins = []
    for l in f.split("\n"):
        m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
    print(f"== {name}: {len(ins)} instructions")
    for a, op, rest in ins:
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t and int(t.group(1), 16) < a:
                tgt = int(t.group(1), 16)
                body = [i for i in ins if tgt <= i[0] <= a]
                if len(body) < 100:
                    continue
                c = collections.Counter(i[1].split(".")[0] if not i[1].startswith("IMAD.") else
                                        ("IMAD.IADD/MOV/SHL" if re.match(r"IMAD\.(IADD|MOV|SHL|U32)", i[1]) else "IMAD")
                                        for i in body)
                print(f"  loop {tgt:#x}..{a:#x}: {len(body)} instructions")
                print("   ", ", ".join(f"{k} {v}" for k, v in sorted(c.items(), key=lambda x: -x[1])))
