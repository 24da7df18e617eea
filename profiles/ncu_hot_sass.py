"""Top SASS instructions by executed count / stall samples from `ncu --page source --csv`."""
import csv, sys, subprocess
f = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(['ncu', '-i', f, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = rows[2:]
ia = hdr.index('Source'); ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
tot = sum(int(r[ie]) for r in data if r[ie].isdigit()); tots = sum(int(r[isamp]) for r in data if r[isamp].isdigit())
print('total warp-instructions', tot, 'samples', tots)
hot = [i for i, r in enumerate(data) if r[ie].isdigit() and int(r[ie]) > 0.004 * tot]
for i in hot[:topn * 4]:
    r = data[i]
    print('%5d %-70s exec %5.2f%%  samples %5.2f%%' % (i, r[ia][:70], 100 * int(r[ie]) / tot, 100 * int(r[isamp]) / max(tots, 1)))
