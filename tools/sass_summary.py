"""Opcode counts per kernel of libsc_b200.so (cuobjdump -sass), written to profiles/."""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "superscreen_b200", "libsc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
kernels, cur, arch = collections.OrderedDict(), None, set()
for line in out.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m: arch.add(m.group(1))
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter()); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
groups = [("DMMA", r"^DMMA"), ("DFMA", r"^DFMA"), ("DADD/DMUL", r"^(DADD|DMUL)"), ("DSETP", r"^DSETP"), ("MUFU.RSQ64H", r"^MUFU\.RSQ64H"),
          ("MUFU.RCP64H", r"^MUFU\.RCP64H"), ("UBLKCP", r"^UBLKCP"), ("SYNCS", r"^SYNCS"), ("LDG", r"^LDG"), ("STG", r"^STG"),
          ("LDS", r"^LDS"), ("STS", r"^STS"), ("SHFL", r"^SHFL"), ("BAR", r"^BAR"), ("ATOM/RED", r"^(ATOM|RED|ATOMG)"),
          ("LDGSTS", r"^LDGSTS"), ("CCTL/PREF", r"^(CCTL|LDGDEPBAR|PREFETCH)"), ("UTMALDG", r"^UTMALDG"), ("UTCxMMA", r"^UTC"), ("LDTM", r"^LDTM"), ("total", r".")]
lines = ["# cuobjdump -sass superscreen_b200/libsc_b200.so : opcode counts per kernel (static instruction counts)",
         f"# cubin architectures: {sorted(arch)}",
         "# fp64 tensor cores on sm_100a are reached through mma.sync.m8n8k4.f64 -> SASS DMMA (there is no fp64 tcgen05 kind:",
         "# UTC*MMA / LDTM / UTMALDG are expected to be 0); TMA appears as 1-D bulk copies cp.async.bulk -> UBLKCP + mbarrier SYNCS",
         "# (the LU operands are pre-packed fragment-major by the panel-solve kernels, so no tensor-map TMA is needed);",
         "# cp.async (the fine-tile latency kernels and the pipelined panel solves of the LU) appears as LDGSTS.",
         "", "kernel".ljust(64) + "".join(g[0].rjust(12) for g in groups)]
tot = collections.Counter()
for name, cnt in kernels.items():
    row = []
    for g, pat in groups:
        v = sum(c for op, c in cnt.items() if re.match(pat, op)); row.append(v); tot[g] += v
    d = re.sub(r"\(.*", "", demangle(name)).replace("void ", "").replace("scb::", "")
    lines.append(d[:63].ljust(64) + "".join(str(v).rjust(12) for v in row))
lines.append("ALL KERNELS".ljust(64) + "".join(str(tot[g[0]]).rjust(12) for g in groups))
path = os.path.join(root, "profiles", sys.argv[1] if len(sys.argv) > 1 else "r02_sass_summary.txt")
open(path, "w").write("\n".join(lines) + "\n")
print(path, len(kernels), "kernels")
