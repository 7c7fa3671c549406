"""Writes profiles/<round>_sass_excerpt.md: the SASS lines that prove which hardware path the hot kernels take
(cuobjdump -sass of the in-tree libminilp_b200.so; runs on the CPU box).
   python scripts/sass_excerpt.py r02"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(ROOT, "minilp_b200", "libminilp_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
        funcs[cur].append(line.strip())


def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]


WANT = {
    "k_price_partial_tma<4>": (r"UBLKCP|SYNCS|ELECT|UTMA", "bulk copy (TMA unit, 1-D cp.async.bulk) + mbarrier ring"),
    "k_price_partial_tma<8>": (r"UBLKCP|SYNCS", "same kernel, 4096-column tiles (narrow shards)"),
    "k_chain_primal": (r"ATOM|RED\.|MEMBAR|LDG\.E\.[0-9.]*STRONG|CCTL|ERRBAR", "grid barriers: fence + atomic arrive + spin"),
    "k_lu_cluster": (r"UCGABAR|CGABAR|MAPA|ST\.E.*CLUSTER|LD\.E.*CLUSTER|BAR\.", "cluster barriers / distributed shared memory"),
    "k_exchange_p2p": (r"MEMBAR\.SC\.SYS|MEMBAR|LDG\.E\.[0-9.]*(SYS|STRONG)|ST\.E\.[0-9.]*(SYS|STRONG)|NANOSLEEP", "system-scope fences and volatile peer loads"),
}
out = [f"# SASS excerpts ({tag}) — `cuobjdump -sass minilp_b200/libminilp_b200.so`, sm_100a", "",
       "Only the instructions that identify the hardware path are listed, with a per-kernel opcode histogram of the arithmetic "
       "and memory instructions.  `-fmad=false`: there must be no DFMA in any kernel (the reference never fuses a*b+c).", ""]
total_dfma = 0
dfma_by = {}
for name, lines in funcs.items():
    dn = demangle(name)
    nd = sum("DFMA" in l for l in lines)
    total_dfma += nd
    if nd:
        dfma_by[dn.replace("void ", "")] = (nd, sum("MUFU.RCP64H" in l or "MUFU.RSQ64H" in l for l in lines))
    for key, (pat, why) in WANT.items():
        if dn.replace("void ", "") == key or (key in dn and "<" not in key):
            hist = collections.Counter()
            for l in lines:
                m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
                if m:
                    hist[m.group(1).split(".")[0]] += 1
            out.append(f"## {dn}  ({len(lines)} instructions) — {why}")
            keep = {k: v for k, v in hist.items() if k in ("DADD", "DMUL", "DFMA", "LDG", "LDS", "STG", "STS", "UBLKCP", "SYNCS", "ATOM",
                                                          "ATOMG", "RED", "MEMBAR", "BAR", "SHFL", "REDUX", "LDGSTS", "DSETP", "MUFU", "UCGABAR_ARV", "UCGABAR_WAIT")}
            out.append("opcode histogram: " + ", ".join(f"{k} {v}" for k, v in sorted(keep.items())))
            out.append("```")
            shown = 0
            for l in lines:
                if re.search(pat, l) and shown < 24:
                    out.append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l))
                    shown += 1
            out.append("```")
            out.append("")
out.append("## DFMA audit")
out.append("`-fmad=false` forbids contracting a*b+c.  The only DFMA left are the Newton steps of the correctly rounded IEEE f64 "
           "DIVISION sequence (each is seeded by MUFU.RCP64H); kernels without a division — every price-out, every tall-skinny "
           "product, the LU trailing update — have none:")
out.append("")
out.append("| kernel | DFMA | MUFU.RCP64H seeds (= divisions) |")
out.append("|---|---|---|")
for k, (nd, nr) in sorted(dfma_by.items()):
    out.append(f"| {k} | {nd} | {nr} |")
nodiv = [demangle(n).replace("void ", "") for n, l in funcs.items() if not any("DFMA" in x for x in l)]
out.append("")
out.append("no DFMA at all: " + ", ".join(sorted(nodiv)))
path = os.path.join(ROOT, "profiles", f"{tag}_sass_excerpt.md")
open(path, "w").write("\n".join(out) + "\n")
print(path, len(out), "lines; DFMA:", total_dfma)
