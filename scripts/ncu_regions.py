"""Instructions / stall samples of an ncu report by source line, with the stall reasons: python scripts/ncu_regions.py rep [file-substr] [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else 'admit.cuh'; top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
blocks = []; cur = None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": cur = {"file": r[1], "hdr": None, "lines": []}; blocks.append(cur)
    elif r[0] == "Line No": cur["hdr"] = r
    elif cur is not None and cur["hdr"] and r[0] != "": cur["lines"].append(r)
for b in blocks:
    if sub not in b['file']: continue
    h = b['hdr']; L = b['lines']
    ie = h.index("Instructions Executed"); si = h.index("Warp Stall Sampling (All Samples)")
    tot = sum(int(l[ie]) for l in L if l[ie].isdigit()); ts = sum(int(l[si]) for l in L if l[si].isdigit())
    if tot < 1000: continue
    print(b['file'], 'instr', tot, 'samples', ts)
    stall_cols = [i for i, n in enumerate(h) if n.startswith('stall_')]
    agg = collections.Counter()
    for l in L:
        for i in stall_cols:
            if l[i].isdigit(): agg[h[i]] += int(l[i])
    print('stall reasons:', ', '.join(f"{k[6:]} {v*100//max(1,sum(agg.values()))}%" for k, v in agg.most_common(8)))
    for l in sorted([l for l in L if l[si].isdigit()], key=lambda l: -int(l[si]))[:top]:
        rs = sorted(((int(l[i]), h[i][6:]) for i in stall_cols if l[i].isdigit() and int(l[i])), reverse=True)[:2]
        print(f"{int(l[si])*100/ts:5.1f}% in {int(l[ie])*100/tot:4.1f}% L{l[0]:>4}: {l[1].strip()[:110]:110s} {rs}")
