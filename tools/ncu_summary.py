"""Turn an ncu report into the markdown summary kept under profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.md "title / command line"."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (realtime; reads low on .2CTA kernels)"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor-memory pipe active % (92 % on cuBLAS bf16 at peak: profiles/r2_tensor_pipe_calibration.md)"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe inst (UTCHMMA/UTMALDG issue)"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes (xbar2l1tex)"),
    ("derived__lts__lts2xbar_bytes.sum.per_second", "L2->xbar rate"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
]


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# %s\n\nSource: `%s` (ncu --set full --clock-control none --import-source on). "
                "Numbers under the profiler are for shares/ratios, never bench values.\n\n" % (title, rep))
        for r in data:
            name = r[idx["Kernel Name"]]
            f.write("## `%s`\n\n| metric | value |\n|---|---|\n" % name.split("(")[0])
            for k, label in KEYS:
                if k in idx:
                    f.write("| %s | %s %s |\n" % (label, r[idx[k]], units[idx[k]]))
            f.write("\n")
    print("wrote", out, len(data), "kernels")


if __name__ == "__main__":
    main()
