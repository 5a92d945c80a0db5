"""Static SASS instruction count per source line of one kernel.  usage: python tools/sass_lines.py <mangled-name-substring> [top]"""
import collections, os, re, subprocess, sys, tempfile
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "wakeword_trainer_home_b200", "lib", "libwwfeat.so")
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", os.path.join(d, cub)], capture_output=True, text=True).stdout
cur, fn, cnt = None, None, collections.Counter()
for l in txt.splitlines():
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m: fn = m.group(1); continue
    if fn is None or sys.argv[1] not in fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l): cnt[cur] += 1
tot = sum(cnt.values()); print("total", tot)
src = {}
for (f, ln), n in cnt.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    path = os.path.join(os.path.dirname(so), "..", "csrc", f)
    line = ""
    if os.path.exists(path):
        L = open(path).read().splitlines(); line = L[ln - 1].strip() if ln <= len(L) else ""
    print(f"{n:5d} {100*n/tot:4.1f}% {f}:{ln}  {line[:100]}")
