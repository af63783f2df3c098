"""Summarise a tools/trace_chunk.py dump: per-item/per-pass segment durations of worker warp 0 (stream 1)."""
import re
import sys

txt = open(sys.argv[1]).read()
streams = re.split(r'--- stream \d+: \d+ events\n', txt)[1:]
ev = [[(int(m.group(1)), int(m.group(2))) for m in re.finditer(r'(\d+)@(\d+)', s)] for s in streams]
w = [e for e in ev[1] if e[0] >= 20]
prev = None
line = []
for tag, c in w[:int(sys.argv[2]) if len(sys.argv) > 2 else 80]:
    line.append(f"{tag}@{c}" + (f"(+{c - prev})" if prev is not None else ""))
    prev = c
print(" ".join(line))
starts = [c for t, c in w if t == 30]
if len(starts) > 1:
    print("item cycles:", [b - a for a, b in zip(starts, starts[1:])])
