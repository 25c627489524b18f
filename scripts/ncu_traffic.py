"""Measure DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum) of the kernels bench.py reports a roofline for,
AT THE BENCH's frames-per-launch, and store them in profiles/r02_traffic.json (bench.py reads that file; nothing is hard-coded there).
    python scripts/ncu_traffic.py [frames_per_launch=32]          (run on the GPU box; needs ncu on PATH)"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
METRICS = "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"


def run(script, args, env=None):
    out = os.path.join(ROOT, "gpurun_out", "_traffic.csv")
    cmd = ["ncu", "--clock-control", "none", "--profile-from-start", "off", "--metrics", METRICS, "--csv", "--log-file", out,
           sys.executable, os.path.join(ROOT, "scripts", script)] + [str(a) for a in args]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env={**os.environ, **(env or {})}, timeout=600)
    rows = list(csv.reader(open(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hi]
    ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    ker = {}
    for r in rows[hi + 1:]:
        if len(r) > vi:
            ker.setdefault(int(r[ii]), {"name": r[ki].split("(")[0].replace("void ", "").replace("lavb::", "")})[r[mi]] = float(r[vi].replace(",", ""))
    return [ker[k] for k in sorted(ker)]


res = {}
for enc in ("tiled", "sorted"):
    ks = run("pillar_layers.py", [B], {"LAVB_PILLAR_ENCODER": enc})
    res["pillar_" + enc] = {"frames": B, "dram_bytes": sum(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"] for k in ks),
                            "us": sum(k["gpu__time_duration.sum"] for k in ks) / 1e3,
                            "kernels": [{"name": k["name"][:48], "us": k["gpu__time_duration.sum"] / 1e3,
                                         "dram_mb": (k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"]) / 1e6} for k in ks],
                            "source": f"ncu dram__bytes_read.sum+dram__bytes_write.sum over the encoder's launches, scripts/pillar_layers.py {B} (120 000 stacked points per frame)"}
ks = run("heads_conv_layer.py", [B])
k = max(ks, key=lambda k: k["gpu__time_duration.sum"])
res["heads_conv"] = {"frames": B, "dram_bytes": k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"], "us": k["gpu__time_duration.sum"] / 1e3,
                     "tensor_pipe_pct": k["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"],
                     "source": f"ncu, conv_umma_kernel 384->256 3x3 @160x160, {B} frames per launch (scripts/heads_conv_layer.py)"}
# every conv_umma_kernel launch of ONE tick of the product pipeline (bench.py's eager pass, B frames in one agent group)
out = os.path.join(ROOT, "gpurun_out", "_traffic_tick.csv")
subprocess.run(["ncu", "--clock-control", "none", "--metrics", METRICS, "--csv", "--log-file", out, "-c", "4000", sys.executable,
                os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--batch", str(B), "--pipelines", "1", "--no-cpu-baseline",
                "--no-train", "--no-gpu-reference", "--no-graphs"], check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=900)
rows = list(csv.reader(open(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hi]
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
ker = {}
for r in rows[hi + 1:]:
    if len(r) > vi:
        ker.setdefault(int(r[ii]), {"name": r[ki]})[r[mi]] = float(r[vi].replace(",", ""))
ids = sorted(ker)
marks = [i for i in ids if "erf_stem" in ker[i]["name"]]
if len(marks) >= 2:
    tick = [ker[i] for i in ids if marks[-2] <= i < marks[-1]]
    um = [k for k in tick if "conv_umma_kernel" in k["name"]]
    res["umma_tick"] = {"frames": B, "launches": len(um), "dram_bytes": sum(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"] for k in um),
                        "us": sum(k["gpu__time_duration.sum"] for k in um) / 1e3, "tick_launches": len(tick),
                        "tick_us": sum(k["gpu__time_duration.sum"] for k in tick) / 1e3,
                        "source": f"ncu over one eager tick of bench.py --batch {B} --pipelines 1 --no-graphs: all conv_umma_kernel launches between two erf_stem launches"}
json.dump(res, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
