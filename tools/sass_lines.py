"""Static SASS instruction counts per source line of one kernel (needs -lineinfo).
  python tools/sass_lines.py <object or cubin> <substring of the mangled kernel name> [min_count]
"""
import collections
import re
import subprocess
import sys
import tempfile
import os


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    minc = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    d = tempfile.mkdtemp()
    cub = obj
    if not obj.endswith(".cubin"):
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, stdout=subprocess.DEVNULL)
        cub = os.path.join(d, [f for f in os.listdir(d) if f.endswith(".cubin")][0])
    txt = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
    cnt, ops, cur, on = collections.Counter(), collections.defaultdict(collections.Counter), None, False
    total = collections.Counter()
    for l in txt.splitlines():
        if l.startswith("//--------------------- .text."):
            on = pat in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m and cur:
            cnt[cur] += 1
            ops[cur][m.group(2)] += 1
            total[m.group(2)] += 1
    for k, v in sorted(cnt.items()):
        if v >= minc:
            print(f"{k[0]}:{k[1]:<5d} {v:5d}  {dict(ops[k].most_common(6))}")
    print("total", sum(total.values()), dict(total.most_common(25)))


if __name__ == "__main__":
    main()
