#!/usr/bin/env python
"""Turn the ncu artefacts in gpurun_out/ into the small, committed summaries under profiles/.

  profiles/<tag>_launches.md    per-kernel share of one step / one decode (from the gpu__time_duration launch list)
  profiles/<tag>_kernels.json   per-launch metrics of the `--set full` captures (duration, DRAM bytes, tensor %, ...)
Run here (no GPU needed): python tools/summarize_profiles.py r1b
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def short(name):
    n = name.split("(")[0]
    for junk in ("void ", "ltx2::<unnamed>::", "unnamed>::", "ltx2::"):
        n = n.replace(junk, "")
    return n.strip()


def launch_list(path, title):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(short(r[kn]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    lines = [f"### {title}", "", f"total kernel time {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches "
             "(ncu `gpu__time_duration.sum`, `--clock-control none`; cold-cache and serialised: compare shares)", "",
             "| kernel | launches | ms | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f} % |")
    return "\n".join(lines) + "\n"


def full_capture(path):
    if path.endswith(".csv"):         # exported on the GPU box by tools/capture_profiles.sh (`ncu -i ... --page raw --csv`)
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": short(r[hdr.index("Kernel Name")])}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[k] = float(r[i].replace(",", ""))
                except ValueError:
                    d[k] = r[i]
                d[k + " [unit]"] = units[i]
        res.append(d)
    return res


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(OUT, exist_ok=True)
    md = [f"# ncu launch lists ({tag})\n"]
    for f, title in ((f"launches_{tag}_dit.csv", "one DiT denoising step (19B, N=3456) -- tools/profile_step.py"),
                     (f"launches_{tag}_dit_fp8.csv", "the same step with the FP8 linear path (fp8_linear) -- tools/profile_step.py --fp8"),
                     (f"launches_{tag}_vae.csv", "one VAE chunk decode (7 latent frames -> 49 frames @ 512x768) -- tools/profile_vae.py")):
        p = os.path.join(SRC, f)
        if os.path.exists(p):
            md.append(launch_list(p, title))
            with open(p) as src, open(os.path.join(OUT, f), "w") as dst:
                dst.write(src.read())
    open(os.path.join(OUT, f"{tag}_launches.md"), "w").write("\n".join(md))
    caps = {}
    for f in sorted(os.listdir(SRC)):
        if f.startswith("prof_") and (f.endswith(f"_{tag}.ncu-rep") or f.endswith(f"_{tag}.csv")):
            caps[f] = full_capture(os.path.join(SRC, f))
    json.dump(caps, open(os.path.join(OUT, f"{tag}_kernels.json"), "w"), indent=1)
    print(open(os.path.join(OUT, f"{tag}_launches.md")).read())
    for f, ks in caps.items():
        for k in ks:
            rd, wr = k.get("dram__bytes_read.sum", 0), k.get("dram__bytes_write.sum", 0)
            print(f, k["kernel"][:40], "us", k.get("gpu__time_duration.sum"), k.get("gpu__time_duration.sum [unit]"),
                  "dram", rd, k.get("dram__bytes_read.sum [unit]"), wr, "tensor%",
                  k.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))


if __name__ == "__main__":
    main()
