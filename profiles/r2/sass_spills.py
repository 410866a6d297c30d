"""Where a kernel's local-memory (spill / stack) accesses sit: STL / LDL instructions of the section by source line.
usage: sass_spills.py ev.sass <mangled-name-fragment>"""
import re, sys, collections
path, frag = sys.argv[1], sys.argv[2]
insec = False; line = 0; c = collections.Counter(); first = {}
for ln in open(path):
    if ln.startswith("\t.section"):
        insec = frag in ln and ".text." in ln; continue
    if not insec: continue
    m = re.search(r'//## File ".*", line (\d+)', ln)
    if m: line = int(m.group(1)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*);', ln)
    if m and re.search(r'\b(STL|LDL)\b', m.group(2)):
        op = 'STL' if 'STL' in m.group(2) else 'LDL'
        c[(line, op)] += 1; first.setdefault(line, int(m.group(1), 16))
for (l, op), n in sorted(c.items()):
    print(f"line {l:5d} {op} x{n:3d}  first at {first[l]:#x}")
