"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        hdr = r; continue
    if hdr and r[0].isdigit() and len(r) == len(hdr):
        iS = hdr.index('Warp Stall Sampling (All Samples)')
        iI = hdr.index('Instructions Executed')
        def num(x):
            try: return int(x)
            except Exception: return 0
        out.append((num(r[iS]), num(r[iI]), cur, int(r[0]), r[1].strip()[:100]))
tot = sum(o[0] for o in out) or 1; toti = sum(o[1] for o in out) or 1
print("samples", tot, "instr", toti)
for s, ins, f, ln, src in sorted(out, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100 * s / tot, 100 * ins / toti, f, ln, src))
