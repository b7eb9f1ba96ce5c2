"""Summarise an event trace of the persistent sampler kernel (SEEME_DP_TRACE=<file>; cluster 0 / rank 0, one step).
roles: 0 producer (1 unit start, 2 A visible, 3 loads issued), 1 MMA (4 first operands landed, 5 last commit issued),
2 epilogue row 64 (6 accumulator ready, 9 slice stored, 10 signalled), 3 statistics exchanges of that thread (7 send, 8 received)."""
import sys
from collections import defaultdict

ev = defaultdict(list)
for line in open(sys.argv[1]):
    r, t, c = line.split()
    ev[int(r)].append((int(t), int(c)))
t0 = min(t for r in ev for t, _ in ev[r])
mhz = float(sys.argv[2]) if len(sys.argv) > 2 else 1965.0
span = max(t for r in ev for t, _ in ev[r]) - t0
print(f"step span: {span} cycles = {span / mhz:.1f} us at {mhz:.0f} MHz")
prod = [(t - t0, c) for t, c in ev[0]]
mma = [(t - t0, c) for t, c in ev[1]]
epi = [(t - t0, c) for t, c in ev[2]]
st = [(t - t0, c) for t, c in ev[3]]
# per stage: acc ready -> signalled (epilogue), signalled -> A visible (exchange), A visible -> first operands (TMA), first operands -> acc ready
acc = [t for t, c in epi if c == 6]
sig = [t for t, c in epi if c == 10]
sto = [t for t, c in epi if c == 9]
vis = [t for t, c in prod if c == 2]
first = [t for t, c in mma if c == 4]
last = [t for t, c in mma if c == 5]
print(f"{len(acc)} stages, {len(vis)} units, {len(st) // 2} statistics exchanges")
sx = [(st[i + 1][0] - st[i][0]) for i in range(0, len(st) - 1, 2)]
if sx:
    print(f"statistics exchange: mean {sum(sx) / len(sx):.0f} cycles, min {min(sx)}, max {max(sx)}, total {sum(sx)} ({100 * sum(sx) / span:.0f}% of the step)")
epi_t = [s - a for a, s in zip(acc, [x for x in sig if x > acc[0]])] if acc else []
tot_epi = 0
for a in acc:
    nxt = [s for s in sig if s > a]
    if nxt:
        tot_epi += nxt[0] - a
print(f"epilogue (acc ready -> signalled), incl. statistics exchanges: total {tot_epi} cycles ({100 * tot_epi / span:.0f}%)")
fence = sum((min(s for s in sig if s > x) - x) for x in sto if any(s > x for s in sig))
print(f"fence + signal after the stores: total {fence} cycles ({100 * fence / span:.0f}%), mean {fence / max(1, len(sto)):.0f}")
# signalled -> next A visible at the producer
tot_x = 0; n_x = 0
for s in sig:
    nxt = [v for v in vis if v > s]
    if nxt:
        tot_x += nxt[0] - s; n_x += 1
print(f"signal -> A visible at the producer: total {tot_x} ({100 * tot_x / span:.0f}%), mean {tot_x / max(1, n_x):.0f}")
tot_t = 0; n_t = 0
for v in vis:
    nxt = [f for f in first if f > v]
    if nxt:
        tot_t += nxt[0] - v; n_t += 1
print(f"A visible -> first operands landed (TMA latency): total {tot_t} ({100 * tot_t / span:.0f}%), mean {tot_t / max(1, n_t):.0f}")
tot_m = 0
for a in acc:
    prev = [f for f in first if f < a]
    # first 'first operands' event after the previous acc
    pa = [x for x in acc if x < a]
    lo = pa[-1] if pa else 0
    cand = [f for f in first if lo < f < a]
    if cand:
        tot_m += a - cand[0]
print(f"first operands -> accumulator ready (MMA + commit): total {tot_m} ({100 * tot_m / span:.0f}%)")
if "-v" in sys.argv:
    allv = sorted([(t, "P", c) for t, c in prod] + [(t, "M", c) for t, c in mma] + [(t, "E", c) for t, c in epi] + [(t, "S", c) for t, c in st])
    for t, r, c in allv[:int(sys.argv[sys.argv.index("-v") + 1]) if len(sys.argv) > sys.argv.index("-v") + 1 else 200]:
        print(t, r, c)
