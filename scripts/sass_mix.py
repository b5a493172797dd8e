#!/usr/bin/env python3
"""Static SASS instruction mix per kernel of the built library (cuobjdump -sass): memory / atomic / async-copy /
FP64 / warp-collective mnemonics - the evidence behind DESIGN.md's statements about which instructions each engine
is made of (no tensor-core, TMA or TMEM instructions: the path is not a contraction)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "net_ensembles_b200", "lib", "libnet_ensembles_cuda.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
groups = [("LDG (global load)", r"^LDG"), ("STG (global store)", r"^STG"), ("LDS (shared load)", r"^LDS"), ("STS (shared store)", r"^STS"),
          ("LDGSTS (cp.async)", r"^LDGSTS"), ("ATOMG/RED (global atomic)", r"^(ATOMG|ATOM\b|RED)"), ("ATOMS (shared atomic)", r"^ATOMS"),
          ("LDL/STL (local: spills)", r"^(LDL|STL)"), ("DADD/DFMA/DMUL (fp64)", r"^(DADD|DFMA|DMUL)"), ("MUFU (rcp seed of fp64 divide)", r"^MUFU"),
          ("SHFL", r"^SHFL"), ("VOTE/MATCH/REDUX (warp collectives)", r"^(VOTE|MATCH|REDUX)"), ("POPC/FLO (bit counts)", r"^(POPC|FLO)"),
          ("BAR/WARPSYNC/UCGABAR (barriers)", r"^(BAR|WARPSYNC|UCGABAR|BSYNC)"), ("IMAD/IADD3/LOP3/SHF (integer)", r"^(IMAD|IADD3|LOP3|SHF|VIADD|LEA|ISETP|SEL|PRMT)"),
          ("tensor core / TMA / TMEM (UTC*MMA, HMMA, UTMALDG, LDTM)", r"^(UTC|HMMA|UTMA|UBLKCP|LDTM|STTM)")]
cur, mix = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        mix[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        mix[cur]["total"] += 1
        for name, pat in groups:
            if re.match(pat, op):
                mix[cur][name] += 1
demangle = subprocess.run(["c++filt"] + list(mix), capture_output=True, text=True).stdout.splitlines()
for mangled, nice in zip(mix, demangle):
    c = mix[mangled]
    if c["total"] < 150:
        continue
    print("== %s  (%d SASS instructions)" % (nice.replace("nec::", "")[:110], c["total"]))
    print("   " + ", ".join("%s %d" % (name.split(" (")[0], c[name]) for name, _ in groups if c[name] or name.startswith("tensor")))
