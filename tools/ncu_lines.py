"""Per-source-line summary of an ncu source page (needs -lineinfo and --import-source on):
    ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --print-source cuda,sass > cs.csv
    python tools/ncu_lines.py cs.csv [top_n]
Prints, for the hottest lines: stall samples, instructions executed, shared / global wavefront columns."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr = None
    files, cur_file = {}, None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr) or r[0] == "":
            continue
        files.setdefault(cur_file, []).append(r)
    col = lambda name: hdr.index(name)
    c_samp, c_inst = col("# Samples"), col("Instructions Executed")
    c_shw = col("L1 Wavefronts Shared")
    c_glob = col("L1 Tag Requests Global")
    c_l2 = col("L2 Theoretical Sectors Global")
    stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    allrows = [(f, r) for f, rs in files.items() for r in rs]
    num = lambda x: float(x) if x not in ("", "-") else 0.0
    tot_s = sum(num(r[c_samp]) for _, r in allrows)
    tot_i = sum(num(r[c_inst]) for _, r in allrows)
    print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
    allrows.sort(key=lambda fr: -num(fr[1][c_samp]))
    print(f"{'file:line':28s} {'samp%':>6s} {'inst%':>6s} {'shWf(M)':>8s} {'gReq(M)':>8s} {'L2sec(M)':>8s}  top stalls | source")
    for f, r in allrows[:top]:
        st = sorted(((num(r[i]), h[6:]) for h, i in stall_cols), reverse=True)[:3]
        sts = " ".join(f"{h}:{v / max(num(r[c_samp]), 1) * 100:.0f}%" for v, h in st if v > 0)
        print(f"{(f + ':' + r[0]):28s} {num(r[c_samp]) / tot_s * 100:6.2f} {num(r[c_inst]) / tot_i * 100:6.2f} "
              f"{num(r[c_shw]) / 1e6:8.2f} {num(r[c_glob]) / 1e6:8.2f} {num(r[c_l2]) / 1e6:8.2f}  {sts:40s} | {r[1].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
