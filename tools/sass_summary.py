"""Static SASS mnemonic counts per kernel from the in-tree objects -> profiles/sass_r01_summary.json."""
import collections, json, re, subprocess, sys, glob, os
OBJ = "physs_gp_b200/csrc/_obj"
WANT = [r"seq_filter_kernel<4, 4, 1, false, false, false>", r"seq_smooth_kernel<4, 4, 0, false, false, true>",
        r"rt_filter_kernel<\d+, \d+, false, \d+, 4, 1>", r"rt_smooth_kernel<\d+, \d+, false, \d+, 4>",
        r"rt_filter_summary_kernel<8, 8, false, 8, 4, 8>", r"rt_smooth_summary_kernel<8, 8, false, 8, 4>",
        r"ps_filter_scan_kernel<8>", r"ps_smooth_scan_kernel<8>", r"ps_filter_scan_reg_kernel<2>", r"ps_smooth_scan_reg_kernel<2>",
        r"kf_vjp_kernel<4, 4, false>", r"cvi_site_kernel<1, 1, 1, 1>", r"pendulum_ell_kernel"]
KEEP = ["DFMA", "DMUL", "DADD", "LDG.E.128", "LDG.E.64", "STG.E.128", "STG.E.64", "LDS.128", "LDS.64", "STS.128", "STS.64",
        "LDGSTS", "UBLKCP", "SYNCS", "MUFU.RCP64H", "MUFU.RSQ64H", "LDL", "STL", "BAR", "WARPSYNC", "SHFL"]
out = {}
for obj in sorted(glob.glob(os.path.join(OBJ, "*.o"))):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = name if any(re.search(w, name) for w in WANT) else None
            if cur:
                out[cur] = collections.Counter()
            continue
        if cur:
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                op = m.group(1)
                out[cur]["total_instructions"] += 1
                for k in KEEP:
                    if op == k or op.startswith(k + ".") or (k in ("LDGSTS", "UBLKCP", "SYNCS", "LDL", "STL", "BAR", "SHFL") and op.startswith(k)):
                        out[cur][k] += 1
                        break
res = {"how": "cuobjdump -sass of the in-tree objects (static instruction counts per kernel, sm_100a); tools/sass_summary.py",
       "kernels": {k.replace("physs::", ""): dict(v) for k, v in sorted(out.items())}}
json.dump(res, open("profiles/sass_r01_summary.json", "w"), indent=1)
print(len(out), "kernels")
for k, v in res["kernels"].items():
    print(k[:90], {x: v[x] for x in ("total_instructions", "DFMA", "LDS.128", "LDG.E.128", "UBLKCP", "LDGSTS", "LDL") if x in v})
