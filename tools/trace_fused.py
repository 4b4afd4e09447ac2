"""Summarise a CAFE_GPU_TRACE dump of the fused K2 kernel (CTA 0): per phase time per (op, half)."""
import sys, collections
import numpy as np
lines = open(sys.argv[1]).read().split('\n')
cta = np.array([list(map(int, l.split()[1:])) for l in lines if l.startswith('cta')])
if len(cta):
    t0 = cta[:,2].min()
    dur = (cta[:,3]-cta[:,2])/1e3
    print("CTAs", len(cta), "duration us: min %.0f mean %.0f max %.0f; kernel span %.0f us" % (dur.min(), dur.mean(), dur.max(), (cta[:,3].max()-t0)/1e3))
    print("start skew us max", (cta[:,2].max()-t0)/1e3)
    per_sm = collections.defaultdict(list)
    for c,(smid,a,b,nmb) in zip(cta[:,0],cta[:,1:]): per_sm[smid].append(((a-t0)/1e3,(b-t0)/1e3,nmb,c))
    ends = {k: max(x[1] for x in v) for k,v in per_sm.items()}
    worst = max(ends, key=ends.get); best = min(ends, key=ends.get)
    print("SM count", len(per_sm), "CTAs per SM", collections.Counter(len(v) for v in per_sm.values()))
    print("worst SM", worst, sorted(per_sm[worst])); print("best SM", best, sorted(per_sm[best]))
    print("m-blocks per SM:", collections.Counter(sum(x[2] for x in v) for v in per_sm.values()))
rows = [list(map(int, l.split())) for l in lines if l and not l.startswith('cta')]
a = np.array(rows)
ev, w, t0, t1, t2, t3, kind = a.T
clk = 1.965e3  # cycles per us
base = t0.min()
print("events", len(set(ev)), "span us", (t3.max() - base) / clk)
for k in (0, 1):
    m = kind == k
    if not m.any():
        continue
    print(f"kind {k}: n={m.sum()//8}  kloop {np.mean(t1[m]-t0[m])/clk:.2f} us  epilogue {np.mean(t2[m]-t1[m])/clk:.2f} us  fence+barrier {np.mean(t3[m]-t2[m])/clk:.2f} us")
# per-warp skew at barrier arrival
for k in (0, 1):
    m = kind == k
    evs = sorted(set(ev[m]))
    sk = []
    for e in evs[:400]:
        mm = ev == e
        sk.append((t2[mm].max() - t2[mm].min()) / clk)
    print(f"kind {k}: arrival skew mean {np.mean(sk):.2f} us  max {np.max(sk):.2f}")
# gap between the end of one event (barrier release) and the start of next
e_sorted = sorted(set(ev))
gaps = []
for e0, e1 in zip(e_sorted, e_sorted[1:]):
    gaps.append((t0[ev == e1].min() - t3[ev == e0].max()) / clk)
print("gap between events us: mean", np.mean(gaps))
m = kind == 1
for e in sorted(set(ev[m]))[:3]:
    mm = ev == e
    print("event", e, [(int(ww), round((a_-base)/clk,1), round((b-base)/clk,1), round((c-base)/clk,1), round((d-base)/clk,1)) for ww,a_,b,c,d in zip(w[mm],t0[mm],t1[mm],t2[mm],t3[mm])])
