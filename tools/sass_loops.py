"""List the loops (backward branches) of one kernel in `cuobjdump -sass` output with an opcode histogram
per loop body.  Usage: python tools/sass_loops.py <sass file> <substring of mangled kernel name> [min body size]"""
import re
import sys
from collections import Counter


def main(path, needle, min_body=20):
    txt = open(path).read()
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n")[0]
        if needle not in name:
            continue
        ins = []
        for line in f.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
        print(name, len(ins), "instructions")
        shown = 0
        for k, (addr, op, rest) in enumerate(ins):
            if shown >= 12:
                print("  ... (more loops not shown)")
                break
            if op.startswith("BRA"):
                t = re.search(r"0x([0-9a-f]+)", rest)
                if t and int(t.group(1), 16) < addr:
                    tgt = int(t.group(1), 16)
                    body = [o for a, o, _ in ins if tgt <= a <= addr]
                    if len(body) >= min_body:
                        c = Counter(o.split(".")[0] if not o.startswith(("STG", "LDS", "LDG", "SHFL")) else o for o in body)
                        print("  loop 0x%x..0x%x: %d instr  %s" % (tgt, addr, len(body), dict(c.most_common(12))))
                        shown += 1


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 20)
