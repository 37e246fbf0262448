"""Join `ncu --page source --csv` (per-SASS-instruction executed counts) with `nvdisasm -g` line info and print the
executed warp instructions per source line (top N).  usage: sass_by_line.py <cubin> <kernel substring> <ncu sass csv>"""
import csv, re, subprocess, sys, collections
cubin, kname, ncsv = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# collect instructions of the kernel in order with their current line annotation
in_k = False; cur = ("?", 0); ins = []
for ln in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
    if m:
        in_k = kname in m.group(1); continue
    if not in_k: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), (m.group(3) or "").split("/")[-1], int(m.group(4) or 0)); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((cur, m.group(2).strip()))
rows = list(csv.reader(open(ncsv)))
hdr = rows[1]
ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples") if "# Samples" in hdr else None
body = rows[2:]
print(f"disasm instrs {len(ins)}, ncu rows {len(body)}", file=sys.stderr)
agg = collections.Counter(); samp = collections.Counter(); ops = collections.defaultdict(collections.Counter)
n = min(len(ins), len(body))
for i in range(n):
    cur, text = ins[i]
    cnt = int(float(body[i][ci] or 0))
    key = cur[:2] if not (len(cur) > 2 and cur[2]) else (cur[0], cur[1], cur[2], cur[3])
    agg[key] += cnt
    if cs is not None: samp[key] += int(float(body[i][cs] or 0))
    ops[key][text.split()[0] if not text.startswith("@") else text.split()[1]] += cnt
tot = sum(agg.values()); ts = sum(samp.values()) or 1
print(f"total warp instr {tot}")
for key, c in agg.most_common(topn):
    top = ", ".join(f"{o}:{v*100//max(c,1)}%" for o, v in ops[key].most_common(4))
    print(f"{100*c/tot:5.1f}% inst {100*samp[key]/ts:5.1f}% samp  {key}  [{top}]")
