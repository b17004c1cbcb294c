"""parse the per-CTA lines of the attention timeline build: cycles per CTA, spread of end times"""
import sys, re, statistics
rows = [tuple(map(int, re.findall(r"\d+", l))) for l in sys.stdin if l.startswith("cta ")]
cyc = [r[3] for r in rows]; end = [r[4] for r in rows]
print(f"{len(rows)} CTAs on {len(set(r[1] for r in rows))} SMs; cycles min {min(cyc)} median {statistics.median(cyc)} max {max(cyc)}; end-time spread {max(end) - min(end)} ns")
by_sm = {}
for r in rows: by_sm.setdefault(r[1], []).append(r[3])
print("CTAs per SM:", sorted(set(len(v) for v in by_sm.values())))
slow = sorted(rows, key=lambda r: -r[3])[:5]; fast = sorted(rows, key=lambda r: r[3])[:5]
print("slowest", slow); print("fastest", fast)
