"""gpurun_out/launches_<tag>.csv (ncu --metrics gpu__time_duration.sum ... --csv) -> profiles/launches_<tag>.{csv,md}.
usage: python scripts/launch_list.py r02 "<command line that was profiled>" """
import collections, csv, shutil, subprocess, sys
tag = sys.argv[1]
cmd = sys.argv[2] if len(sys.argv) > 2 else ""
src = f"gpurun_out/launches_{tag}.csv"
rows = [r for r in csv.reader(open(src)) if r]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[ix["Metric Unit"]]]
    k = r[ix["Kernel Name"]]
    e = agg.setdefault(k, [0, 0.0, r[ix["Grid Size"]], r[ix["Block Size"]]])
    e[0] += 1; e[1] += v
tot = sum(e[1] for e in agg.values())
own = sum(e[1] for k, e in agg.items() if "dcgru::" in k)
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
out = [f"# ncu launch list ({tag}, commit {commit}) -- `{cmd}`",
       "# (cold-cache, serialised launches: compare SHARES, not absolute times)",
       f"# total device time of the captured launches: {tot:.3f} ms", "",
       "| kernel | launches | total ms | share | grid | block |", "|---|---:|---:|---:|---|---|"]
for k, e in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    out.append(f"| `{k[:70]}` | {e[0]} | {e[1]:.3f} | {100 * e[1] / tot:.1f}% | {e[2]} | {e[3]} |")
out += ["", f"own kernels (dcgru::*): {100 * own / tot:.1f}% of the captured device time; the rest are torch glue kernels "
        "(loss, gather, copies, NCCL)."]
open(f"profiles/launches_{tag}.md", "w").write("\n".join(out) + "\n")
shutil.copy(src, f"profiles/launches_{tag}.csv")
print("\n".join(out[:16]))
