#!/usr/bin/env python
"""Where a kernel's warps wait: reads `ncu -i X.ncu-rep --page source --csv` (SASS view, needs --import-source on / -lineinfo is
not required) and prints, per kernel in the report, the share of warp-stall samples by stall reason and the top instructions.
  ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_hotspots.py src.csv [top_n]
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    rows = list(csv.reader(open(path)))
    # the CSV holds one table per kernel: a header row containing 'Address' starts each
    tables, cur, name = [], None, "?"
    for r in rows:
        if "Address" in r and "Source" in r:
            cur = {"hdr": r, "rows": [], "name": name}
            tables.append(cur)
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
        elif r and cur is None:
            name = " ".join(r)[:100]
        elif r and len(r) < 4:
            name = " ".join(r)[:100]
            cur = None
    seen = set()
    for t in tables:
        key = (t['name'], len(t['rows']), tuple(t['rows'][0][:3]) if t['rows'] else ())
        if key in seen:
            continue
        seen.add(key)
        idx = {h: i for i, h in enumerate(t["hdr"])}
        samp = idx.get("Warp Stall Sampling (All Samples)")
        if samp is None:
            continue
        reasons = [h for h in t["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(float(r[samp] or 0) for r in t["rows"])
        if tot == 0:
            continue
        by_reason = collections.Counter()
        for r in t["rows"]:
            for h in reasons:
                by_reason[h] += float(r[idx[h]] or 0)
        rs = sum(by_reason.values()) or 1.0
        print(f"== {t['name']}  ({int(tot)} samples, {len(t['rows'])} SASS instructions)")
        print("   stall reasons: " + ", ".join(f"{k[6:]} {100 * v / rs:.0f}%" for k, v in by_reason.most_common(7)))
        ranked = sorted(t["rows"], key=lambda r: -float(r[samp] or 0))[:top_n]
        for r in ranked:
            v = float(r[samp] or 0)
            why = max(reasons, key=lambda h: float(r[idx[h]] or 0))
            print(f"   {100 * v / tot:5.1f}%  {why[6:]:<14} {r[idx['Source']].strip()[:90]}")


if __name__ == "__main__":
    main()
