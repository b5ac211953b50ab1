#!/usr/bin/env python
"""Per-barrier-region digest of one kernel's SASS source page: executed warp-instructions, stall samples, shared-memory
wavefronts (and the excess from bank conflicts) between consecutive BAR.SYNC instructions.
  ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv ; python tools/ncu_regions.py src.csv [label ...]
Labels name the regions in order (the part after the last barrier gets the last label)."""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    labels = sys.argv[2:]
    hdr, data = rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    ia, isrc, iex, ismp = col["Address"], col["Source"], col["Instructions Executed"], col["# Samples"]
    iw, ie = col["L1 Wavefronts Shared"], col["L1 Wavefronts Shared Excessive"]
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = int(data[0][ia], 16)
    num = lambda r, i: int(r[i] or 0)
    tot_ex, tot_s, tot_w = sum(num(r, iex) for r in data), sum(num(r, ismp) for r in data), sum(num(r, iw) for r in data)
    print(f"# {rows[0][1] if len(rows[0]) > 1 else ''}")
    print(f"# {len(data)} SASS instructions, {tot_ex} warp-instructions executed, {tot_s} stall samples, {tot_w} shared-memory wavefronts")
    all_st = {hdr[i][6:]: sum(num(r, i) for r in data) for i in stalls}
    ss = sum(all_st.values())
    print("# stall reasons: " + ", ".join(f"{k} {100 * v / ss:.0f}%" for k, v in sorted(all_st.items(), key=lambda x: -x[1])[:8]))
    regions, cur = [], []
    for r in data:
        cur.append(r)
        if "BAR.SYNC" in r[isrc]:
            regions.append(cur); cur = []
    regions.append(cur)
    print(f"{'region':34s} {'SASS':>5s} {'exec %':>7s} {'samples %':>9s} {'smem wavefronts %':>18s} {'excess %':>8s}  top stalls")
    for k, reg in enumerate(regions):
        ex, s, w, e = (sum(num(r, i) for r in reg) for i in (iex, ismp, iw, ie))
        st = {hdr[i][6:]: sum(num(r, i) for r in reg) for i in stalls}
        sst = max(sum(st.values()), 1)
        top = ", ".join(f"{n} {100 * v / sst:.0f}%" for n, v in sorted(st.items(), key=lambda x: -x[1])[:3])
        name = labels[k] if k < len(labels) else f"region {k}"
        lo, hi = int(reg[0][ia], 16) - base, int(reg[-1][ia], 16) - base
        print(f"{name[:34]:34s} {len(reg):5d} {100 * ex / tot_ex:7.1f} {100 * s / tot_s:9.1f} {100 * w / max(tot_w, 1):18.1f} {100 * e / max(w, 1):8.1f}  {top}   [{lo:#x}..{hi:#x}]")


if __name__ == "__main__":
    main()
