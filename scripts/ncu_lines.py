"""Per-source-line instruction counts of one kernel from an .ncu-rep (read here, no GPU needed):
   python scripts/ncu_lines.py file.ncu-rep [top N]   ->  line, warp instructions, share, avg lanes, stall samples, source text"""
import csv
import io
import subprocess
import sys


def main(rep, top=60):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    lines = []
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "":
            continue
        ie = hdr.index("Instructions Executed")
        te = hdr.index("Thread Instructions Executed")
        ss = hdr.index("# Samples")
        try:
            lines.append((int(r[0]), int(r[ie]), int(r[te]), int(r[ss]), r[1]))
        except ValueError:
            pass
    tot = sum(l[1] for l in lines)
    tots = sum(l[3] for l in lines)
    print("total warp instructions %d, samples %d" % (tot, tots))
    acc = 0
    for ln, ie, te, ss, src in sorted(lines, key=lambda l: -l[1])[:top]:
        acc += ie
        print("%5d %12d %5.2f%% (cum %5.1f%%) lanes %5.1f samples %5.2f%%  %s" %
              (ln, ie, 100.0 * ie / tot, 100.0 * acc / tot, te / max(ie, 1), 100.0 * ss / max(tots, 1), src.strip()[:110]))
    return lines


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60)
