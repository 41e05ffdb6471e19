"""Top CUDA source lines by warp-stall samples of an ncu report captured with --import-source on:
python scripts/ncu_hotspots.py gpurun_out/final_admit.ncu-rep 'regex:k_admit' > profiles/r1_ncu_k_admit_source_hotspots.txt"""
import csv, io, subprocess, sys

rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = {"file": r[1], "fn": None, "hdr": None, "lines": []}; blocks.append(cur)
    elif r[0] == "Function Name":
        cur["fn"] = r[1]
    elif r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and r[0] != "":
        cur["lines"].append(r)
for b in blocks:
    h = b["hdr"]; si = h.index("Warp Stall Sampling (All Samples)"); ie = h.index("Instructions Executed")
    rows = [l for l in b["lines"] if l[si].isdigit()]
    tot = sum(int(l[si]) for l in rows)
    if tot < 500 or "grove" not in b["file"]:
        continue
    print(f"== {b['fn']}\n   file {b['file']}, {tot} stall samples")
    for l in sorted(rows, key=lambda l: -int(l[si]))[:25]:
        print(f"  {int(l[si]) * 100 / tot:5.1f}%  warp-instr {l[ie]:>9}  L{l[0]:>4}: {l[1].strip()[:130]}")
