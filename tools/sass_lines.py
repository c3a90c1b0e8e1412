#!/usr/bin/env python
"""Static SASS instruction count per CUDA source line.
usage: nvdisasm -c -g <cubin> > dis.txt; python tools/sass_lines.py dis.txt <function-substring> [topN]"""
import collections
import re
import sys

path, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
cur_fn, cur = None, None
cnt = collections.Counter()
for l in open(path, errors="ignore"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        cur_fn = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if cur_fn and want in cur_fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        cnt[cur] += 1
tot = sum(cnt.values())
print("total SASS instructions:", tot, f"({tot * 16 / 1024:.0f} KiB)")
cache = {}
regions = collections.Counter()
for (f, ln), c in cnt.items():
    regions[(f.split("/")[-1], ln // 50 * 50)] += c
print("--- by 50-line region")
for (f, r), c in sorted(regions.items(), key=lambda kv: -kv[1])[:25]:
    print(f"{f}:{r}-{r+49}  {c}  {100*c/tot:.1f}%")
print("--- by line")
for (f, ln), c in cnt.most_common(top):
    if f not in cache:
        try:
            cache[f] = open(f).read().split("\n")
        except OSError:
            cache[f] = []
    t = cache[f][ln - 1].strip()[:90] if ln <= len(cache[f]) else ""
    print(f"{f.split('/')[-1]}:{ln}  {c}  {100*c/tot:.1f}% | {t}")
