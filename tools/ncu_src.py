import csv, sys, collections
path, kern = sys.argv[1], sys.argv[2]
rows = csv.reader(open(path))
cur_file = cur_fn = None; hdr = None
by_line = collections.Counter(); samp_line = collections.Counter(); by_op = collections.Counter(); samp_op = collections.Counter()
src_text = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed"); continue
    if kern not in (cur_fn or ""): continue
    try:
        ins = float(r[iI] or 0); smp = float(r[iS] or 0)
    except Exception: continue
    if r[0] != "":   # CUDA source line row
        key = (cur_file, int(r[0])); by_line[key] += ins; samp_line[key] += smp; src_text[key] = r[1].strip()[:90]
    else:           # SASS row
        op = r[3].split()[0] if r[3] else "?"
        if op.startswith("@"): op = r[3].split()[1]
        op = op.rstrip(';')
        by_op[op] += ins; samp_op[op] += smp
tot = sum(by_op.values()) or 1; ts = sum(samp_op.values()) or 1
print("total inst", tot, "samples", ts)
print("--- opcodes")
for op, n in by_op.most_common(25): print(f"{op:28s} {n/tot*100:6.2f}% inst  {samp_op[op]/ts*100:6.2f}% samples")
tl = sum(by_line.values()) or 1; tsl = sum(samp_line.values()) or 1
print("--- lines by samples")
for k, n in samp_line.most_common(40): print(f"{k[0]}:{k[1]:<5d} {by_line[k]/tl*100:6.2f}% inst {n/tsl*100:6.2f}% samp | {src_text[k]}")
