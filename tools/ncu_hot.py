"""Hot basic blocks of the first kernel in an .ncu-rep (SASS page): executed warp-instructions per block,
opcode mix and dominant stall.  usage: ncu_hot.py rep [top]"""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 15
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ix = {k: i for i, k in enumerate(hdr)}
ins = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"): break
    ins.append(r)
tot = sum(int(r[ix["Instructions Executed"]]) for r in ins)
samples = sum(int(r[ix["# Samples"]]) for r in ins)
print(f"{len(ins)} SASS instructions, {tot} warp-instructions executed, {samples} samples")
# blocks = maximal runs with identical exec count
blocks = []
cur = None
for n, r in enumerate(ins):
    e = int(r[ix["Instructions Executed"]])
    if cur is None or cur["e"] != e:
        cur = {"e": e, "lo": n, "hi": n, "rows": []}
        blocks.append(cur)
    cur["hi"] = n; cur["rows"].append(r)
def opc(s):
    s = re.sub(r"^@!?U?P\w+\s+", "", s.strip())
    return s.split()[0].split(".")[0]
stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
for b in sorted(blocks, key=lambda b: -b["e"] * len(b["rows"]))[:top]:
    n = len(b["rows"]); w = b["e"] * n
    mix = collections.Counter(opc(r[ix["Source"]]) for r in b["rows"])
    st = collections.Counter()
    smp = 0
    for r in b["rows"]:
        smp += int(r[ix["# Samples"]])
        for k in stall_cols:
            st[k] += int(r[ix[k]] or 0)
    print(f"[{b['lo']:5d}-{b['hi']:5d}] n={n:4d} exec={b['e']:9d} share={100*w/tot:5.1f}% samples={100*smp/max(samples,1):5.1f}%  "
          + " ".join(f"{k}:{v}" for k, v in mix.most_common(8)) + "  | " + " ".join(f"{k[6:]}:{v}" for k, v in st.most_common(3)))
