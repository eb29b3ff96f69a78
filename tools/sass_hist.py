"""Per-opcode and per-barrier-segment executed-instruction histogram of one kernel from an ncu report:
   python tools/sass_hist.py <prof.ncu-rep> <kernel regex> [launch index]"""
import collections, csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name": cur = []; blocks.append(cur); continue
    if r and r[0] == "Address": hdr = r; continue
    if cur is not None and len(r) > 5: cur.append(r)
b = blocks[idx]
ii, si, ss = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ii]) for r in b)
ops, samp = collections.Counter(), collections.Counter()
for r in b:
    t = r[si].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += int(r[ii]); samp[op] += int(r[ss])
print("launches captured:", len(blocks), " total warp instr:", tot)
for op, n in ops.most_common(24): print(f"   {op:12s} {n:10d} {n/tot*100:5.1f}%  samples {samp[op]}")
seg = acc = sam = 0
for r in b:
    acc += int(r[ii]); sam += int(r[ss])
    if "BAR.SYNC" in r[si] or "EXIT" in r[si]:
        if acc > tot * 0.002: print(f"segment {seg}: instr {acc:10d} ({acc/tot*100:4.1f}%) samples {sam}")
        seg += 1; acc = sam = 0
