"""Summarise an `ncu --page source --csv` dump: the SASS instructions where warps stall most."""
import csv
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[2:]:
        try:
            s = float(r[ci["# Samples"]])
        except (ValueError, IndexError):
            continue
        top_stall = max(stalls, key=lambda k: float(r[ci[k]] or 0))
        data.append((s, r[ci["Source"]].strip()[:90], top_stall, r[ci["Instructions Executed"]]))
    tot = sum(d[0] for d in data) or 1.0
    print(f"total samples {tot:.0f}, {len(data)} SASS instructions")
    for s, text, st, ex in sorted(data, reverse=True)[:top]:
        print(f"{100 * s / tot:5.1f}%  {st:22s} exec={ex:>9s}  {text}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
