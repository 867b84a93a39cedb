"""Per-source-line warp-stall samples of one kernel from an `ncu --set full --import-source on` report.

usage: python profiles/source_hotspots.py <report.ncu-rep> <kernel regex> [launch index] [top N]
Reads `ncu -i <rep> --page source --csv --print-source cuda,sass`, keeps the per-line rows (the ones with a line number),
and prints the lines holding the most samples together with their dominant stall reasons."""
import csv
import io
import subprocess
import sys


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    idx = sys.argv[3] if len(sys.argv) > 3 else "1"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id",
                          f"::regex:{kre}:{idx}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    fname, hdr, recs = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif len(r) > 10 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            recs.append((fname, r))
    if not hdr:
        print("no correlated rows"); return
    col = {h: i for i, h in enumerate(hdr)}
    s_i = col["# Samples"]
    stall_cols = [(h, i) for h, i in col.items() if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[s_i] or 0) for _, r in recs)
    print(f"total samples {tot}; lines with samples {sum(1 for _, r in recs if int(r[s_i] or 0))}")
    recs.sort(key=lambda fr: -int(fr[1][s_i] or 0))
    for f, r in recs[:top]:
        n = int(r[s_i] or 0)
        if not n:
            break
        st = sorted(((int(r[i] or 0), h[6:]) for h, i in stall_cols), reverse=True)[:3]
        st = " ".join(f"{h}:{v}" for v, h in st if v)
        print(f"{n:6d} {100.0 * n / tot:5.1f}%  {f}:{r[0]:>4s}  {st:45s} | {r[1].strip()[:110]}")


if __name__ == "__main__":
    main()
