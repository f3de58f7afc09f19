"""SASS evidence for the fused evaluation kernel (run here, no GPU): mnemonic counts that prove the Blackwell-native path
(UTCHMMA.2CTA = tcgen05.mma cta_group::2, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS =
mbarrier) and an excerpt of the MMA issue loop.   python scripts/sass_summary.py > profiles/r02_sass_fused_kernel.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "iwae_b200", "lib", "libiwae_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        funcs[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln).rstrip())
name = next(f for f in funcs if "iwae_eval_tc_kernelILb1ELb0ELb0ELb0E" in f)
body = funcs[name]
print(f"# {so.replace(ROOT + '/', '')}  (cuobjdump -sass), kernel {name}")
print(f"# = tc::iwae_eval_tc_kernel<FAST=true, TRACE=false, INJECT=false, TERMS=false>: the instantiation bench.py times; {len(body)} instructions")
cnt = collections.Counter()
for ln in body:
    m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m: cnt[m.group(1)] += 1
keys = ["UTCHMMA.2CTA", "UTCBAR.2CTA.MULTICAST", "LDTM.x16", "UBLKCP.S.G", "UTCATOMSWS", "MUFU.TANH", "MUFU.EX2", "MUFU.LG2", "MUFU.SIN", "MUFU.COS", "MUFU.SQRT", "F2FP.SATFINITE.F16.F32.PACK_AB"]
print("# mnemonic counts (static):")
for k in keys + sorted(k for k in cnt if k.startswith(("SYNCS", "UTC", "LDTM", "STTM", "UBLK", "UTMA")) and k not in keys):
    if cnt.get(k): print(f"#   {k:40s} {cnt[k]}")
print("# all kernels of the library and their tcgen05 / bulk-copy instruction counts:")
for f, b in funcs.items():
    c = sum(1 for l in b if "UTCHMMA" in l); l = sum(1 for x in b if "LDTM" in x); u = sum(1 for x in b if "UBLKCP" in x)
    if c or l or u: print(f"#   {f[:110]:110s} UTCHMMA {c:3d}  LDTM {l:3d}  UBLKCP {u:3d}")
i = next(n for n, l in enumerate(body) if "UTCHMMA" in l)
print("# --- excerpt: the MMA issue loop (one slab = up to four cta_group::2 MMAs, then the multicast commit) ---")
for l in body[max(0, i - 28): i + 75]: print(l)
