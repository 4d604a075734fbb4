"""Rank CUDA source lines of one kernel by warp-stall samples.
    ncu -i REPORT.ncu-rep --page source --print-source cuda,sass --csv > src.csv
    python profiles/ncu_source_lines.py src.csv [top]
(the report must have been captured with --set full --import-source on from a -lineinfo build)"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr, cur_file, agg = None, None, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "":
            continue                      # SASS rows carry no line number
        try:
            n = int(r[6])
        except ValueError:
            continue
        stall = {hdr[i][6:]: int(r[i] or 0) for i in range(len(hdr))
                 if hdr[i].startswith("stall_") and "Not Issued" not in hdr[i]}
        key = (cur_file, int(r[0]))
        if key in agg:
            agg[key][0] += n
            for k, v in stall.items():
                agg[key][2][k] = agg[key][2].get(k, 0) + v
        else:
            agg[key] = [n, r[1], stall]
    tot = sum(v[0] for v in agg.values())
    print("total samples", tot)
    for (f, l), (n, src, st) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        best = sorted(st.items(), key=lambda x: -x[1])[:3]
        print("%6d %5.1f%% %s:%-5d %-70s %s" % (n, 100.0 * n / tot, f[:14], l, src.strip()[:70], best))
    tots = {}
    for v in agg.values():
        for k, c in v[2].items():
            tots[k] = tots.get(k, 0) + c
    print("stall reasons:", sorted(tots.items(), key=lambda x: -x[1]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
