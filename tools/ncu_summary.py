"""Summaries of the ncu outputs a gpurun call brought back (development tool; writes markdown to stdout).

    python tools/ncu_summary.py launches gpurun_out/r01d_launches.csv
    python tools/ncu_summary.py full gpurun_out/r01d_tsolve.ncu-rep
"""
import csv, io, re, subprocess, sys


def short(name):
    m = re.match(r"void (?:mmpgo::)?(?:<?unnamed>::)?(\w+)(<[^>]*>)?", name)
    if not m:
        return name[:60]
    t = (m.group(2) or "").replace("(int)", "").replace("mmpgo::", "")
    return m.group(1) + t


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    h = rows[0]
    kn, bs, gs, mv = h.index("Kernel Name"), h.index("Block Size"), h.index("Grid Size"), h.index("Metric Value")
    seq = [(short(r[kn]), r[bs], r[gs], float(r[mv]) / 1e3) for r in rows[1:]]
    # one step = from the k_prox launch of one iteration to the k_prox launch of the next (bench.py --steps 2
    # --warmup 1: the second and third k_prox; later k_prox launches belong to the profile passes)
    prox = [i for i, s in enumerate(seq) if s[0].startswith("k_prox")]
    lo, hi, nsteps = prox[1], prox[2], 1
    agg = {}
    for name, b, g, us in seq[lo:hi]:
        a = agg.setdefault((name, b, g), [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(v[1] for v in agg.values())
    print("| kernel | block | grid | launches/step | ms/step | share | avg us |")
    print("|---|---|---|---:|---:|---:|---:|")
    for (name, b, g), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %s | %s | %.1f | %.3f | %.3f | %.1f |" % (name, b, g, n / nsteps, us / nsteps / 1e3, us / tot, us / n))
    print("\nSum of kernel time: %.2f ms/step under ncu (%d launches/step)." % (tot / nsteps / 1e3, sum(v[0] for v in agg.values()) / nsteps))
    ts = sum(v[1] for k, v in agg.items() if k[0].startswith("k_tsolve") or k[0].startswith("k_mf_solve"))
    print("Share of the translation solve (k_mf_solve, or k_tsolve + k_tsolve_lite on the PCG path): %.3f" % (ts / tot))


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[0]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]
    units = rows[1]
    for r in rows[2:]:
        print("\n### `%s`\n\n| metric | value |\n|---|---|" % short(r[h.index("Kernel Name")]))
        vals = {}
        for k in want:
            if k in h:
                vals[k] = (r[h.index(k)], units[h.index(k)])
                print("| %s | %s %s |" % (k, r[h.index(k)], units[h.index(k)]))
        def tob(k):
            v, u = vals[k]
            return float(v.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
        def tos(k):
            v, u = vals[k]
            return float(v.replace(",", "")) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1e-9)
        tb = tob("dram__bytes_read.sum") + tob("dram__bytes_write.sum")
        print("| **DRAM traffic per launch** | %.3f GB (%.0f GB/s over the launch) |" % (tb / 1e9, tb / tos("gpu__time_duration.sum") / 1e9))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
