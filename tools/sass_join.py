#!/usr/bin/env python
"""Joins ncu's per-SASS-instruction counters with nvdisasm line info: executed warp instructions and
stall samples per source line and per opcode (used for profiles/*_source_hotspots.txt).

    python tools/sass_join.py REPORT.ncu-rep LIB.so KERNEL_MANGLED_SUBSTRING [--top N] [--op OPCODE]
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

def main():
    rep, lib, kern = sys.argv[1:4]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    opf = sys.argv[sys.argv.index("--op") + 1] if "--op" in sys.argv else None
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    lines = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"): continue
        txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, on = None, False
        for ln in txt.splitlines():
            if ln.startswith(".text."):
                on = kern in ln
                continue
            if not on: continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
            m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', ln)
            if m: lines[int(m.group(1), 16)] = cur
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]; data = [r for r in rows[2:] if len(r) > 10]
    iA, iS, iN, iAd = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Address")
    base = int(data[0][iAd], 16)
    ex, sm, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    tot = sum(int(r[iA] or 0) for r in data); ts = sum(int(r[iN] or 0) for r in data)
    for r in data:
        off = int(r[iAd], 16) - base
        s = re.sub(r'^@!?U?P\w+\s+', '', r[iS].strip()); op = s.split()[0] if s else "?"
        b = op.split(".")[0]
        if b == "IMAD" and "MOV" in op: b = "IMAD.MOV"
        if opf and b != opf: continue
        k = lines.get(off)
        ex[k] += int(r[iA] or 0); sm[k] += int(r[iN] or 0); ops[k][b] += int(r[iA] or 0)
    src = {}
    print(f"total warp instr {tot}  samples {ts}")
    for k, v in ex.most_common(top):
        text = ""
        if k:
            if k[0] not in src:
                for d in ("rossby-wave-ray-tracing_b200/csrc",):
                    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", d, k[0])
                    src[k[0]] = open(p).read().splitlines() if os.path.exists(p) else []
            if src[k[0]] and k[1] <= len(src[k[0]]): text = src[k[0]][k[1] - 1].strip()[:70]
        o = " ".join(f"{a}:{100*b/tot:.2f}" for a, b in ops[k].most_common(3))
        print(f"{(k[0]+':'+str(k[1])) if k else '?':26s} {100*v/tot:6.2f} {100*sm[k]/max(ts,1):6.2f}  {text:70s} | {o}")

if __name__ == "__main__":
    main()
