"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep (SASS page) joined with
nvdisasm line info of the built .so.   usage: ncu_lines.py report.ncu-rep 'forward_kernelILi128ELi512ELi10' [top]"""
import csv, io, os, re, subprocess, sys, tempfile

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("HMC_B200_LIB") or os.path.join(root, "hmc_b200", "libhmc_b200.so")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
inst = []
for r in rows[2:]:
    if len(r) > ii and r[ia].startswith("0x"):
        inst.append((int(r[ia], 16), int(r[ii] or 0), int(r[isamp] or 0), {k: int(r[c] or 0) for k, c in stall_cols.items()}, r[1].strip()))
base = inst[0][0]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# locate the function section
lines = dis.splitlines()
cur_line = None; in_fn = False; off2line = {}
for ln in lines:
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        in_fn = pat in m.group(1)
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        off2line[int(m.group(1), 16)] = cur_line
agg = {}
tot_i = sum(x[1] for x in inst); tot_s = sum(x[2] for x in inst)
for addr, n, s, st, txt in inst:
    key = off2line.get(addr - base)
    a = agg.setdefault(key, [0, 0, {}])
    a[0] += n; a[1] += s
    for k, v in st.items():
        a[2][k] = a[2].get(k, 0) + v
src_cache = {}
def src(key):
    if not key: return "?"
    f, l = key
    p = os.path.join(root, "hmc_b200", "csrc", f)
    if p not in src_cache:
        src_cache[p] = open(p).read().splitlines() if os.path.exists(p) else []
    return src_cache[p][l - 1].strip()[:100] if l - 1 < len(src_cache[p]) else ""
print(f"total warp-inst {tot_i}  samples {tot_s}")
for key, (n, s, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    top_st = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{100*n/tot_i:5.1f}% inst {100*s/max(tot_s,1):5.1f}% smp  {key[0] if key else '?'}:{key[1] if key else 0:4d}  [{' '.join(f'{k[6:]}={v}' for k,v in top_st)}]  {src(key)}")
