"""Aggregate an ncu source page (--print-source cuda,sass --csv) per CUDA source line.

  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME > src.csv
  python profiles/ncu_lines.py src.csv [top]
"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    fname, hdr, out = "", None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Name":
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 8 and r[0] == "Line No":
            hdr = r
            ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) < len(hdr) or not r[0]:
            continue
        try:
            out.append((int(r[ci]), int(r[cs]), fname, r[0], r[1].strip()[:100]))
        except ValueError:
            pass
    tot, tots = sum(o[0] for o in out), sum(o[1] for o in out)
    print("total warp-instructions %d, samples %d" % (tot, tots))
    for o in sorted(out, reverse=True)[:top]:
        print("%6.2f%% inst %6.2f%% samp  %s:%s  %s" % (100.0 * o[0] / tot, 100.0 * o[1] / max(tots, 1), o[2], o[3], o[4]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
