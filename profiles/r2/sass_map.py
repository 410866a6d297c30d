"""Code-layout map of a kernel from `nvdisasm -g -c` output: for every 2 KB of SASS, the kernel-body source lines that live
there (diagnostic for instruction-cache pressure).  usage: sass_map.py ev.sass <mangled-name-fragment> <first-body-line>"""
import re, sys, collections
path, frag, body0 = sys.argv[1], sys.argv[2], int(sys.argv[3])
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
insec = False; line = 0; rows = []
for ln in open(path):
    if ln.startswith("\t.section"):
        insec = frag in ln and ".text." in ln
        continue
    if not insec: continue
    m = re.search(r'//## File ".*", line (\d+)', ln)
    if m: line = int(m.group(1)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*);', ln)
    if m: rows.append((int(m.group(1), 16), line, m.group(2).strip()))
print("instructions", len(rows), "bytes", rows[-1][0] + 16)
by = collections.defaultdict(list)
for a, l, t in rows: by[a // chunk].append((l, t))
for c in sorted(by):
    ls = [l for l, _ in by[c]]
    body = [l for l in ls if l >= body0]
    inl = collections.Counter(l for l in ls if l < body0)
    ops = collections.Counter(t.split()[0].lstrip('@!P0123456789').split('.')[0] if not t.startswith('@') else t.split()[1].split('.')[0] for _, t in by[c])
    print(f"{c*chunk:6d}: body {min(body) if body else 0:4d}-{max(body) if body else 0:4d} n={len(body):3d} | inl {dict(inl.most_common(4))} | {dict(ops.most_common(4))}")
