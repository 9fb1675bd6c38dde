"""Summarise `ncu -i <rep> --page source --csv` of the fine kernels: per kernel the stall-reason totals over all SASS
instructions, the instructions with the most warp samples, and a coarse map of the kernel in buckets of 200 instructions
(share of samples, memory / MUFU / TMEM / mbarrier instruction counts, dominant stall reasons).
usage: python profiles/src_hotspots.py <source.csv>"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
print("# SYNCS.PHASECHK...TRYWAIT + BRA = mbarrier try-wait (epilogue warps waiting for their tile's MMAs, control warps waiting for "
      "operands / ring stages); EXIT / barrier = the idle control warps")
for si in range(len(secs) - 1):
    hdr = rows[secs[si] + 1]
    ia, isrc, ins = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
    stall = [(h[6:], i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = rows[secs[si] + 2: secs[si + 1]]
    tot = sum(int(r[ins] or 0) for r in body) or 1
    print(f"===== {rows[secs[si]][1][:60]}  instructions {len(body)} samples {tot}")
    st = {}
    for r in body:
        for h, i in stall:
            st[h] = st.get(h, 0) + int(r[i] or 0)
    print("stall totals (% of samples):", {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda x: -x[1])[:10]})
    top = sorted(body, key=lambda r: -int(r[ins] or 0))[:14]
    for r in top:
        s3 = sorted(((h, int(r[i] or 0)) for h, i in stall), key=lambda x: -x[1])[:3]
        print(f"{100 * int(r[ins] or 0) / tot:6.2f}% {r[ia][-5:]} {r[isrc][:60]:60s} {[t for t in s3 if t[1] > 0]}")
    print("bucket  share  LDG STG LDTM MUFU SYNCS  dominant stalls (% of all samples)")
    B = 200
    for b in range(0, len(body), B):
        seg = body[b:b + B]
        n = sum(int(r[ins] or 0) for r in seg)
        if n < tot * 0.005:
            continue
        sb = {}
        for r in seg:
            for h, i in stall:
                sb[h] = sb.get(h, 0) + int(r[i] or 0)
        ops = [(r[isrc].split()[1] if r[isrc].startswith("@") else r[isrc].split()[0]) if r[isrc].split() else "" for r in seg]
        cnt = lambda p: sum(1 for o in ops if o.startswith(p))
        print(f"{b:6d} {100 * n / tot:5.1f}% {cnt('LDG'):4d}{cnt('STG'):4d}{cnt('LDTM'):5d}{cnt('MUFU'):5d}{cnt('SYNCS'):6d}  "
              f"{[(k, round(100 * v / tot, 1)) for k, v in sorted(sb.items(), key=lambda x: -x[1])[:4]]}")
