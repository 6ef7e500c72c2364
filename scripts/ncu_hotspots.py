"""Source-line hot spots of one kernel from an ncu report captured with --import-source on:
    python scripts/ncu_hotspots.py gpurun_out/prof_K.ncu-rep [top]  ->  markdown on stdout
Aggregates the `--page source --print-source cuda,sass` view per CUDA source line (stall
samples, executed warp instructions) so that a profile summary can name file:line."""
import csv
import io
import subprocess
import sys


def hotspots(rep, top=15):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    path, header, lines = None, None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            path = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            header = r
            continue
        if header is None or len(r) < len(header) or r[2] != "-":      # SASS rows carry an address
            continue
        rec = dict(zip(header, r))
        try:
            samples = int(rec["# Samples"])
            insts = int(rec["Instructions Executed"])
        except (KeyError, ValueError):
            continue
        if samples == 0 and insts == 0:
            continue
        key = (path, int(r[0]))
        cur = lines.setdefault(key, [r[1].strip(), 0, 0, {}])
        cur[1] += samples
        cur[2] += insts
        for k, v in rec.items():
            if k.startswith("stall_") and "Not Issued" not in k:
                try:
                    cur[3][k] = cur[3].get(k, 0) + int(v)
                except ValueError:
                    pass
    total_s = sum(v[1] for v in lines.values()) or 1
    total_i = sum(v[2] for v in lines.values()) or 1
    print(f"| file:line | stall samples | warp instructions | top stall | source |\n|---|---:|---:|---|---|")
    for (p, ln), (src, s, i, st) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
        reason = max(st.items(), key=lambda kv: kv[1])[0] if st else "-"
        print(f"| {p}:{ln} | {100.0 * s / total_s:.1f}% | {100.0 * i / total_i:.1f}% | {reason} | `{src[:90]}` |")


if __name__ == "__main__":
    hotspots(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 15)
