#!/usr/bin/env python3
"""Static SASS per source line of one kernel (nvdisasm --print-line-info on the library's cubin): where the spills
(LDL/STL) and the instruction bulk of a kernel sit.  usage: sass_lines.py <mangled-name-substring> [top_n]"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "net_ensembles_b200", "lib", "libnet_ensembles_cuda.so")
want = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.startswith("nec_api.") and f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cub)], capture_output=True, text=True).stdout
cur_fn, cur_line = None, None
spills, total = collections.Counter(), collections.Counter()
for line in txt.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line) or re.match(r"\s*\.section\s+\.text\.(\S+),", line)
    if m:
        cur_fn = m.group(1)
        continue
    if cur_fn is None or want not in cur_fn:
        continue
    m = re.search(r'//## File ".*?/([^/"]+)", line (\d+)', line)
    if m:
        cur_line = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        total[cur_line] += 1
        if m.group(1).startswith(("LDL", "STL")):
            spills[cur_line] += 1
print("%d SASS instructions, %d LDL/STL" % (sum(total.values()), sum(spills.values())))
src_cache = {}
def text(k):
    if k is None: return ""
    f = os.path.join(ROOT, "net_ensembles_b200", "csrc", k[0])
    if f not in src_cache:
        src_cache[f] = open(f).read().splitlines() if os.path.exists(f) else []
    return src_cache[f][k[1] - 1].strip()[:100] if k[1] - 1 < len(src_cache[f]) else ""
print("-- spills by line")
for k, v in spills.most_common(topn):
    print("%4d %s:%s  %s" % (v, k[0] if k else "?", k[1] if k else "?", text(k)))
print("-- instructions by line")
for k, v in total.most_common(topn):
    print("%4d %s:%s  %s" % (v, k[0] if k else "?", k[1] if k else "?", text(k)))
