"""Turn gpurun_out/*.ncu-rep and launches.csv into small tracked summaries under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches_X.md "title"
    python scripts/summarize_ncu.py kernel gpurun_out/prof_K.ncu-rep profiles/r01_K.md "title"
"""
import collections
import csv
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
       "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
       "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
       "launch__occupancy_limit_registers", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum"]


def launches(src, dst, title):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        a = agg.setdefault(r["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` launch list "
                f"({len(rows)} launches, {tot/1e3:.0f} us total; cold-cache, serialised: compare SHARES).\n\n")
        f.write("| share | launches | avg us | kernel |\n|---:|---:|---:|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
            f.write(f"| {v[1]/tot*100:.1f}% | {v[0]} | {v[1]/v[0]/1e3:.1f} | `{k[:110]}` |\n")
        ours = sum(v[1] for k, v in agg.items() if "dgcnn::" in k)
        f.write(f"\nOur kernels (`dgcnn::*`): {ours/tot*100:.1f}% of the listed device time.\n")


def kernel(src, dst, title):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on`, from `{src}`.\n\n")
        for r in rows[2:4]:
            f.write(f"## {r[hdr.index('Kernel Name')][:80]} (launch id {r[0]})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in RAW:
                if m in hdr:
                    f.write(f"| {m} | {r[hdr.index(m)]} | {units[hdr.index(m)]} |\n")
            f.write("\n")
        sass = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--print-source", "sass"],
                              capture_output=True, text=True).stdout
        srows = list(csv.reader(sass.splitlines()))
        h = srows[1]
        idx = {n: i for i, n in enumerate(h)}
        stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        tot = collections.Counter()
        mn = collections.Counter()
        for r in srows[2:]:
            if len(r) < len(h):
                continue
            for s in stalls:
                try:
                    tot[s] += int(r[idx[s]])
                except ValueError:
                    pass
            try:
                op = r[idx["Source"]].split()[0 if not r[idx["Source"]].strip().startswith("@") else 1].split(".")[0]
                mn[op] += int(r[idx["Instructions Executed"]])
            except (ValueError, IndexError):
                pass
        t = sum(tot.values()) or 1
        f.write("## warp stall samples (all launches in the report)\n\n| reason | share |\n|---|---:|\n")
        for s, v in tot.most_common(8):
            f.write(f"| {s} | {v/t*100:.1f}% |\n")
        ti = sum(mn.values()) or 1
        f.write("\n## executed warp instructions by SASS opcode\n\n| opcode | share |\n|---|---:|\n")
        for s, v in mn.most_common(14):
            f.write(f"| {s} | {v/ti*100:.1f}% |\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](*sys.argv[2:5])
