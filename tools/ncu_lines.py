"""Per-source-line instruction counts of one kernel from an .ncu-rep captured with --import-source on:
    python tools/ncu_lines.py rep.ncu-rep kernel_regex [top]"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{kern}"],
                         capture_output=True, text=True).stdout.splitlines()
    rows, hdr, fname, files = [], None, "", set()
    for r in csv.reader(out):
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            if fname in files:
                break                                     # the same file again: the next launch
            files.add(fname)
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0].strip().isdigit():
            r[1] = fname + ": " + r[1].strip()
            rows.append(r)
    ie = hdr.index("Instructions Executed")
    sm = hdr.index("# Samples")
    tot = sum(int(r[ie]) for r in rows)
    tots = sum(int(r[sm]) for r in rows) or 1
    print(f"total warp instructions {tot}, samples {tots}")
    for r in sorted(rows, key=lambda r: -int(r[ie]))[:top]:
        print(f"{int(r[ie]):>10} {100 * int(r[ie]) / tot:5.1f}%  samples {100 * int(r[sm]) / tots:5.1f}%  L{r[0]:>4}  {r[1].strip()[:110]}")


if __name__ == "__main__":
    main()
