"""Per-SASS-instruction summary of one kernel from an ncu report's source page.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --print-source sass > k.csv; python ncu_sass.py k.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if r and r[0].startswith("0x")]
tot_inst = sum(int(r[ix['Instructions Executed']]) for r in data)
tot_s = sum(int(r[ix['# Samples']]) for r in data)
print("total warp-instr %d samples %d sass lines %d" % (tot_inst, tot_s, len(data)))
for k, r in enumerate(data):
    ie = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
    print("%4d %-72s %8d %6d %8s %8s" % (k, r[ix['Source']].strip()[:72], ie // 1000, s, r[ix['L1 Wavefronts Shared']], r[ix['L1 Wavefronts Shared Ideal']]))
