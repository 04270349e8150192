"""Turn an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X.csv) into the markdown table kept under
profiles/.  Usage: python tools/launch_list_summary.py gpurun_out/launches.csv profiles/out.md "title / command line"."""
import collections
import csv
import re
import sys


def main():
    src, out, title = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    rows = [r for r in csv.reader(open(src)) if len(r) > 14]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
        ns = float(r[ix["Metric Value"]].replace(",", ""))
        if r[ix["Metric Unit"]] in ("us", "usecond"):
            ns *= 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values()) or 1.0
    with open(out, "w") as f:
        f.write("# %s\n\nSource: `%s`.  Per-launch times under the profiler are cold-cache and serialised (no PDL overlap): compare "
                "SHARES with the live CUDA-event shares of `bench.py`, never the absolute values.\n\n" % (title, src))
        f.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (name, n, ns / 1e3, ns / 1e3 / n, 100 * ns / total))
    print("wrote", out, len(agg), "kernels")


if __name__ == "__main__":
    main()
