"""Per-source-line summary of an ncu report: executed warp instructions, stall samples and the top stall reasons.
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top_n]
Reads `ncu -i REP --page source --csv --print-source cuda,sass` (needs -lineinfo + --import-source on at capture time)."""
import csv
import io
import subprocess
import sys


def main(rep, top=40):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    cur_file, hdr, lines = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
            d = dict(zip(hdr, r))
            try:
                inst = int(d["Instructions Executed"]); samp = int(d["# Samples"])
            except ValueError:
                continue
            stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
            lines.append((cur_file, int(r[0]), r[1].strip()[:110], inst, samp, stalls))
    tot_i = sum(l[3] for l in lines) or 1
    tot_s = sum(l[4] for l in lines) or 1
    print(f"total warp-instructions {tot_i}, samples {tot_s}")
    agg = {}
    for l in lines:
        for k, v in l[5].items():
            agg[k] = agg.get(k, 0) + v
    print("stall mix:", ", ".join(f"{k} {100*v/tot_s:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    print("| file:line | inst % | samples % | top stalls | source |")
    print("|---|---|---|---|---|")
    for f, ln, src, inst, samp, st in sorted(lines, key=lambda l: -l[4])[:top]:
        ts = " ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"| {f}:{ln} | {100*inst/tot_i:.1f} | {100*samp/tot_s:.1f} | {ts} | `{src}` |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
