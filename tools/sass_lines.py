"""SASS instruction count per source line of one kernel (code-size breakdown; needs -lineinfo).
    python tools/sass_lines.py build/depth_filter.o update_seeds"""
import collections, os, re, subprocess, sys, tempfile
obj, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    sass = subprocess.run(["/usr/local/cuda/bin/nvdisasm", "-g", "-c", cub], cwd=d, capture_output=True, text=True).stdout
cur, fn = None, None
cnt, per_fn = collections.Counter(), collections.Counter()
for l in sass.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        fn = m.group(1)
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
        per_fn[fn] += 1
        if fn and pat in fn:
            cnt[cur] += 1
print("sections:", [(k[-50:], v) for k, v in per_fn.most_common(6)])
tot = sum(cnt.values())
print("total instructions in sections matching", pat, ":", tot, "=", tot * 16 // 1024, "KB")
byfile = collections.Counter()
for (f, ln), c in cnt.items():
    byfile[f] += c
print(byfile.most_common(8))
for k, c in cnt.most_common(top):
    print(f"{c:6d}  {k[0]}:{k[1]}")
