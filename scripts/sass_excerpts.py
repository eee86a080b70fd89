"""profiles/r2_final_sass_excerpts.txt: tensor-core / TMA / TMEM mnemonics per contraction kernel of the built library.
    python scripts/sass_excerpts.py > profiles/r2_final_sass_excerpts.txt"""
import collections, glob, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = glob.glob(os.path.join(ROOT, "mintime-*", "libmintime_b200.so"))[0]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
mangled = re.findall(r"Function : (\S+)", sass)
names = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True).stdout.splitlines()
blocks = re.split(r"\n\s*Function : ", sass)[1:]
SHOW = ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM")
COUNT = SHOW + ("UTCBAR", "HMMA", "LDSM", "FFMA2", "MUFU.TANH", "HMUL2", "LD.E", "ST.E", "LDS", "STS")
print("# cuobjdump -sass libmintime_b200.so at the end of round 2: tensor-core / TMA / TMEM mnemonics per kernel (count) and their first\n"
      "# occurrence.  UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG / UTMASTG / UTMAREDG = cp.async.bulk.tensor load / store /\n"
      "# reduce-add, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, HMMA = mma.sync, LDSM = ldmatrix; LD.E / ST.E = generic loads / stores (the GEMM\n"
      "# kernels had them on SHARED memory until the alignment fix of this round, DESIGN 4d; what is left are global accesses).\n"
      "# Regenerate: python scripts/sass_excerpts.py\n")
for nm, b in zip(names, blocks):
    nm = nm.replace("(anonymous namespace)::", "").replace("mt::", "")
    short = re.sub(r"^void ", "", nm)
    short = short[:short.index(">(") + 1] if ">(" in short else short.split("(")[0]
    if not any(k in short for k in ("gemm_tc", "fused_attn", "mbconv_front", "stem_tc", "attn_group_bwd_mma", "attn_space_mma", "attn_time_mma")):
        continue
    cnt, first = collections.Counter(), {}
    for ln in b.splitlines():
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
        if not m:
            continue
        op = m.group(1)
        for k in COUNT:
            if op.startswith(k):
                key = op if k in SHOW else k
                cnt[key] += 1
                if k in SHOW and key not in first:
                    first[key] = re.sub(r"\s*/\*[^*]*\*/\s*$", "", ln.strip())
                break
    if not any(k.startswith(("UTCHMMA", "HMMA")) for k in cnt):
        continue
    print("## " + short)
    print("   " + ", ".join(f"{k} x{v}" for k, v in sorted(cnt.items())))
    for v in first.values():
        print("      " + v)
    print()
