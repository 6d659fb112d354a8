import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
for i, h in enumerate(hdr):
    if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
        try:
            v = float(vals[i])
        except ValueError:
            continue
        if v > 500:
            print(h.replace('smsp__pcsamp_warps_issue_stalled_', ''), v)
for k in ['sm__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active']:
    print(k, vals[hdr.index(k)])
