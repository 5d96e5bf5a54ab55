"""Stall samples of an ncu capture per CUDA source line (inlined call chains collapsed to the outermost line in the kernel file).

    python tools/ncu_lines.py <report.ncu-rep> <object.o> <mangled-kernel-substring> [top]

Reads the SASS page of the report (`ncu --page source --csv`) and the line table of the object (`nvdisasm -g`), matches
the two instruction streams by position, and prints the samples per source line plus a per-opcode summary.  No GPU needed."""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if "# Samples" in r)
data = rows[rows.index(hdr) + 1:]
iS, iN = hdr.index("Source"), hdr.index("# Samples")
sass = [(r[iS].strip(), int(r[iN] or 0)) for r in data if len(r) > iN]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
lines, cur = [], None
for l in dis[start + 1:]:
    if l.startswith("//---") or l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        # with inlining nvdisasm prints the innermost location first, then "inlined at" lines: keep the outermost one
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.search(r'inlined at "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
if len(lines) != len(sass):
    print(f"warning: {len(lines)} instructions in the object, {len(sass)} in the report", file=sys.stderr)
tot = sum(n for _, n in sass)
per = collections.Counter()
for (s, n), loc in zip(sass, lines):
    per[loc] += n
print(f"total samples {tot}")
src_cache = {}
for loc, n in per.most_common(top):
    text = ""
    if loc:
        path = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", loc[0])
        if os.path.exists(path):
            src_cache.setdefault(path, open(path).read().splitlines())
            text = src_cache[path][loc[1] - 1].strip()[:100]
    print(f"{n:8d} {100 * n / tot:5.1f}%  {loc}  {text}")
