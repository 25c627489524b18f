"""Turn ncu outputs into the markdown tables kept under profiles/.

  python scripts/summarize_ncu.py launches <launch_list.csv[.gz]> [--pass-marker rgb_norm]
      per-kernel time table of the LAST complete pass between two launches of the marker kernel (one perception pass of
      bench.py starts with rgb_norm_kernel); without a marker, the whole list.
  python scripts/summarize_ncu.py rep <capture.ncu-rep>
      key metrics of every kernel in a --set full capture, its instruction mix and the hottest SASS lines by stall samples
      (needs `ncu` on PATH to export the raw / source pages).
"""
import collections, csv, gzip, io, re, subprocess, sys

KEY = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
       "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
       "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def _open(path):
    return io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)


def short(k):
    k = re.sub(r"\(.*", "", k).replace("void ", "").replace("lavb::", "")
    return re.sub(r"\(bool\)|\(int\)", "", k)[:72]


def launches(path, marker=None):
    rows = list(csv.reader(_open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hi]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    L = [(r[ki], float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)) for r in rows[hi + 1:] if len(r) > vi]
    if marker:
        st = [i for i, (k, _) in enumerate(L) if marker in k]
        if len(st) >= 2:
            L = L[st[-2]:st[-1]]
    agg = collections.OrderedDict()
    for k, v in L:
        c = agg.setdefault(short(k), [0, 0.0])
        c[0] += 1
        c[1] += v
    tot = sum(v for _, v in L)
    print(f"{len(L)} launches, {tot:.1f} us summed kernel time\n")
    print("| kernel | launches | us | share |\n|---|---|---|---|")
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {n} | {v:.1f} | {100 * v / tot:.1f} % |")


def rep(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[0]
    for r in rows[2:]:
        d = dict(zip(h, r))
        print(f"### {short(d.get('Kernel Name', '?'))}\n\n| metric | value |\n|---|---|")
        for k in KEY:
            if d.get(k) not in (None, ""):
                print(f"| `{k}` | {d[k]} {rows[1][h.index(k)]} |")
        st = sorted(((float(v or 0), k) for k, v in d.items() if "pcsamp_warps_issue_stalled" in k and not k.endswith("_not_issued")), reverse=True)
        tot = sum(v for v, _ in st) or 1.0
        print("\nstall samples: " + ", ".join(f"{k.split('stalled_')[1]} {100 * v / tot:.0f} %" for v, k in st[:7]) + "\n")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next((i for i, r in enumerate(rows) if r and r[0] == "Address"), None)
    if hi is None:
        return
    h = rows[hi]
    isrc, ismp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    data = []
    for r in rows[hi + 1:]:
        if len(r) == len(h):
            try:
                data.append((r[isrc], float(r[ismp] or 0), float(r[iex] or 0)))
            except ValueError:
                pass
    ts, ti = sum(d[1] for d in data) or 1.0, sum(d[2] for d in data) or 1.0
    mix, smp = collections.Counter(), collections.Counter()
    for s, a, b in data:
        op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
        mix[op] += b
        smp[op] += a
    print("instruction mix (executed / stall samples): " + ", ".join(f"{op} {100 * v / ti:.1f} % / {100 * smp[op] / ts:.1f} %" for op, v in mix.most_common(10)))
    print("\nhottest SASS lines by samples:\n")
    for s, a, b in sorted(data, key=lambda d: -d[1])[:12]:
        print(f"    {100 * a / ts:5.1f} %  x{b:10.0f}  {s[:100]}")


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] not in ("launches", "rep"):
        sys.exit(__doc__)
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[sys.argv.index("--pass-marker") + 1] if "--pass-marker" in sys.argv else None)
    else:
        rep(sys.argv[2])
