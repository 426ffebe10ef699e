"""Per-CUDA-source-line warp-stall samples of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys

def main(path, top=45):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    lines = []
    fname = ""
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-":      # a CUDA line aggregate (Address == "-")
            lines.append((fname, r))
    i_s = hdr.index("Warp Stall Sampling (All Samples)")
    i_ex = hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[i_s] or 0) for _, r in lines)
    print("total samples", tot)
    agg = {hdr[i]: sum(int(r[i] or 0) for _, r in lines) for i in stall}
    print("  ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
    for f, r in sorted(lines, key=lambda fr: -int(fr[1][i_s] or 0))[:top]:
        rs = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall), reverse=True)[:2]
        print(f"{int(r[i_s]):6d} {100*int(r[i_s])/tot:5.1f}%  inst {r[i_ex]:>9s}  {f}:{r[0]:>4s}  {r[1][:90]}   {rs}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
