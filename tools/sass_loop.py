#!/usr/bin/env python
"""Static view of the integrator's stage loop: python tools/sass_loop.py LIB.so|OBJ.o [kernel substring] [--dump]
Finds the innermost backward branch whose body holds the 20 LDS.128 of the staged corner points and prints the
instruction count of that body with an opcode histogram (what round 2 used to compare kernel variants before
spending GPU time: 706 -> 673 -> 639 -> ... instructions)."""
import collections, re, subprocess, sys


def sass(path, kern):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, res = None, {}
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            res[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur:
            res[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return {k: v for k, v in res.items() if kern in k}


def main():
    path = sys.argv[1]
    kern = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "fast19rk45_integrate_coopILi1E"
    for name, ins in sass(path, kern).items():
        best = None
        for a, t in ins:
            m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                lo = int(m.group(1), 16)
                body = [(x, y) for x, y in ins if lo <= x <= a]
                if sum("LDS.128" in y for _, y in body) >= 20 and (best is None or len(body) < len(best)):
                    best = body
        if not best:
            print(name, "no stage loop found")
            continue
        h = collections.Counter()
        for _, t in best:
            op = re.sub(r"^@!?U?P\w+\s+", "", t).split()[0]
            b = op.split(".")[0]
            if b == "IMAD" and "MOV" in op: b = "IMAD.MOV"
            h[b] += 1
        fp64 = sum(v for k, v in h.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
        print(f"{name}: stage loop {len(best)} instructions ({best[0][0]:#x}..{best[-1][0]:#x}), FP64 {fp64}")
        print("  " + "  ".join(f"{k}:{v}" for k, v in h.most_common()))
        if "--dump" in sys.argv:
            for a, t in best: print(f"{a:05x} {t}")


if __name__ == "__main__":
    main()
