"""Decodes the scheduling control fields (stall count, write / read barrier, wait mask) of a SASS range from
`nvdisasm -g -c -hex` output — to see which instruction waits on which outstanding memory operation.
usage: sass_ctrl.py file.sass <mangled-name-fragment> <first-source-line> <last-source-line> [max-rows]"""
import re, sys
path, frag, l0, l1 = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
insec = False; line = 0; rows = []; pend = None
for ln in open(path):
    if ln.startswith("\t.section"):
        insec = frag in ln and ".text." in ln; continue
    if not insec: continue
    m = re.search(r'//## File ".*", line (\d+)', ln)
    if m: line = int(m.group(1)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/', ln)
    if m:
        pend = [int(m.group(1), 16), line, m.group(2).strip(), int(m.group(3), 16)]; continue
    m = re.match(r'\s+/\* (0x[0-9a-f]{16}) \*/', ln)
    if m and pend:
        hi = int(m.group(1), 16)
        ctrl = hi >> 41                                   # bits 105.. of the 128-bit word
        stall, yld, wr, rd, wait = ctrl & 15, (ctrl >> 4) & 1, (ctrl >> 5) & 7, (ctrl >> 8) & 7, (ctrl >> 11) & 63
        rows.append((*pend[:3], stall, wr, rd, wait)); pend = None
on = False; n = 0
for a, l, t, stall, wr, rd, wait in rows:
    if l0 <= l <= l1: on = True
    if on:
        w = "".join(str(i) for i in range(6) if wait >> i & 1)
        print(f"{a:#07x} L{l:<5d} st{stall:<2d} {'W'+str(wr) if wr != 7 else '  '} {'R'+str(rd) if rd != 7 else '  '} {'wait[' + w + ']' if w else '':10s} {t}")
        n += 1
        if n >= (int(sys.argv[5]) if len(sys.argv) > 5 else 120): break
