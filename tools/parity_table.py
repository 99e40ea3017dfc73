"""Tabulates gpurun_out/parity_table.jsonl (written by the GPU parity tests, tests/parity.py::record) as markdown.

    python tools/parity_table.py gpurun_out/parity_table.jsonl > profiles/r2_parity_table.md
"""
import json
import sys


def main(path):
    rows = {}
    for line in open(path):
        r = json.loads(line)
        rows[r["case"]] = r           # last run of a case wins
    print("# Parity table (GPU vs reference fixtures / C oracle), one row per comparison of `pytest -m gpu`\n")
    print("`non-grazing`: live rays on which the reference (fixtures) or the oracle (samples) is insensitive to +-1 ulp moves of "
          "the source -- the set the tolerance is asserted on.  `grazing within tol`: rays OUTSIDE that set (the reference itself is "
          "chaotic there) that nevertheless agree to < 1e-6 deg over the whole trajectory -- reported, not asserted.\n")
    print("| case | slots | live | non-grazing | outliers | max dev on non-grazing [deg] | worst rel. state | grazing | grazing within tol | steps GPU | steps ref |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for case, r in rows.items():
        if "live" not in r:
            continue
        grazing = r["live"] - r["nongrazing"] - r.get("outliers", 0)
        print(f"| {case} | {r['slots']} | {r['live']} | {r['nongrazing']} | {r.get('outliers', 0)} | {r['max_dev_deg_nongrazing']:.2e} | "
              f"{r['worst_rel_state']:.1e} | {grazing} | {r['grazing_within_tol']} | {r['steps_total']} | {r['steps_total_ref']} |")
    print()
    for case, r in rows.items():
        if "live" not in r:
            print(f"* {case}: " + ", ".join(f"{k} = {v}" for k, v in r.items() if k != "case"))
        elif "status_counts_nongrazing" in r:
            print(f"* {case}: died among non-grazing rays {r['died_nongrazing']} (all live: {r['died_all_live']}); status counts on the "
                  f"non-grazing set (1 alive, 2 pole, 3 speed cut-off, 4 NaN state) {r['status_counts_nongrazing']}; died-mask agreement "
                  f"over ALL live rays {r['died_mask_agreement_all_live']:.4f}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/parity_table.jsonl")
