"""A/B of library builds of the same ABI on the bench workload (one subprocess per variant: RWRT_LIB is read at import).

    python rossby-wave-ray-tracing_b200/build.py --variant variants/lib_x.so -DFOO=1      # here (nvcc cross-compiles)
    python tools/ab_variants.py variants/lib_base.so variants/lib_x.so ...               # on the GPU box
Prints per variant: steps/s, ms per pass, ms of the integrate kernel alone, ms of a pass that keeps the ray records.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    wl = os.environ.get("RWRT_BENCH_WORKLOAD", "c2")
    for lib in sys.argv[1:]:
        env = dict(os.environ, RWRT_LIB=os.path.abspath(lib))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "2", "--no-e2e", "--no-cpu",
                            "--no-strong", "--workload", wl], env=env, capture_output=True, text=True)
        try:
            l = json.loads(r.stdout.strip().splitlines()[-1])
            rec = l.get("records") or {}
            print(f"{os.path.basename(lib):28s} {l['value']:.4e} steps/s  pass {l['ms_per_step']:7.2f} ms  integrate {l['roofline']['kernel_ms']:7.2f} ms  "
                  f"frac {l['roofline']['frac']:.4f}  records pass {rec.get('ms_per_step', float('nan')):7.2f} ms", flush=True)
        except Exception as e:
            print(f"{os.path.basename(lib):28s} FAILED: {e}\n{r.stderr[-600:]}", flush=True)


if __name__ == "__main__":
    main()
