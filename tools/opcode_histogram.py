#!/usr/bin/env python
"""Dynamic opcode histogram of one kernel from an ncu report's per-instruction counters (profiles/*_opcode_histogram.txt).

    python tools/opcode_histogram.py REPORT.ncu-rep [kernel regex] > profiles/..._opcode_histogram.txt

Executed warp instructions, stall samples and shared-memory wavefronts by opcode, the per-reason stall totals, and the
instructions that carry shared-memory wavefronts (the `ncu -i ... --page source --csv --print-source sass` page)."""
import collections, csv, io, re, subprocess, sys


def main():
    rep = sys.argv[1]
    kern = sys.argv[2] if len(sys.argv) > 2 else "rk45_integrate_coop"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    idx = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    rows = rows[idx[0]:(idx[1] if len(idx) > 1 else len(rows))]
    name, hdr = rows[0][1], rows[1]
    data = [r for r in rows[2:] if len(r) > 10]
    ix = {h: i for i, h in enumerate(hdr)}
    g = lambda r, k: int(r[ix[k]] or 0)
    tot, ts = sum(g(r, "Instructions Executed") for r in data), sum(g(r, "# Samples") for r in data)
    W = sum(g(r, "L1 Wavefronts Shared") for r in data)
    ex, sm, wf = collections.Counter(), collections.Counter(), collections.Counter()
    for r in data:
        s = re.sub(r"^@!?U?P\w+\s+", "", r[ix["Source"]].strip())
        op = s.split()[0] if s else "?"
        b = op.split(".")[0]
        ex[b] += g(r, "Instructions Executed"); sm[b] += g(r, "# Samples"); wf[b] += g(r, "L1 Wavefronts Shared")
    print(f"# {name}: executed warp instructions by opcode ({rep})")
    print(f"total warp instructions {tot}   ({tot / 592 / 1e6:.1f} M per SM sub-partition)   stall samples {ts}   shared-memory wavefronts {W}")
    fp = sum(ex[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
    print(f"FP64 pipe instructions (DFMA DMUL DADD DSETP): {100 * fp / tot:.1f} %")
    print(f"{'opcode':12s} {'% instr':>8s} {'% stall samples':>16s} {'shared-memory wavefronts':>26s}")
    for b, v in ex.most_common(32):
        print(f"{b:12s} {100 * v / tot:8.2f} {100 * sm[b] / max(ts, 1):16.2f} {wf[b]:26d}")
    print("\nstall reasons (warp samples, all):")
    st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for h, v in sorted(((h, sum(g(r, h) for r in data)) for h in st), key=lambda x: -x[1]):
        if v: print(f"  {h:26s} {100 * v / max(ts, 1):6.2f} %")
    print("\ninstructions with shared-memory wavefronts (executed, wavefronts, ideal, excessive, L1 tag requests):")
    base = int(data[0][ix["Address"]], 16)
    agg = collections.OrderedDict()
    for r in data:
        w = g(r, "L1 Wavefronts Shared")
        if w < W / 2000: continue
        op = re.sub(r"^@!?U?P\w+\s+", "", r[ix["Source"]].strip()).split()[0]
        a = agg.setdefault(op, [0, 0, 0, 0, 0, 0])
        for i, k in enumerate(("Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L1 Wavefronts Shared Excessive",
                               "L1 Tag Requests Global")):
            a[i + 1] += g(r, k)
        a[0] += 1
    for op, a in agg.items():
        print(f"  {op:14s} x{a[0]:<3d} executed {a[1]:12d}  wavefronts {a[2]:12d} ({a[2] / max(a[1], 1):.2f} per instr)  ideal {a[3]:12d}  excessive {a[4]:12d}  tags {a[5]:12d}")


if __name__ == "__main__":
    main()
