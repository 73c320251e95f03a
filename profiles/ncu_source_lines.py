"""Per-source-line hot spots of one kernel from an .ncu-rep captured with --import-source on:
    python profiles/ncu_source_lines.py gpurun_out/prof.ncu-rep <kernel regex> [top N]"""
import csv
import subprocess
import sys


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    fname, idx, lines = "", None, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 5 and r[0] == "Line No":
            idx = {n: i for i, n in enumerate(r)}
            continue
        if idx is None or len(r) <= idx["Instructions Executed"] or r[2] != "-":
            continue          # keep the per-line aggregate rows (Address == "-")
        key = (fname, r[0])
        s, i = num(r[idx["# Samples"]]), num(r[idx["Instructions Executed"]])
        old = lines.get(key, (0.0, 0.0, r[1]))
        lines[key] = (old[0] + s, old[1] + i, r[1])
    ts = sum(v[0] for v in lines.values()) or 1.0
    ti = sum(v[1] for v in lines.values()) or 1.0
    print(f"total samples {ts:.0f}, warp instructions {ti:.0f}")
    for (f, ln), (s, i, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * s / ts:5.1f}% samples {100 * i / ti:5.1f}% inst  {f}:{ln:>4}  {src.strip()[:100]}")


if __name__ == "__main__":
    main()
