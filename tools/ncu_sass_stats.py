#!/usr/bin/env python
"""Opcode mix and the most-stalled instructions of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass`
(the capture must have been taken with --import-source on / --set full). Usage: tools/ncu_sass_stats.py dump.csv [n_samples]
n_samples (default 65536 * 4096) turns instruction counts into warp instructions per processed sample."""
import collections
import csv
import sys


def main(path, nsamp):
    rows = list(csv.reader(open(path)))
    hdr, data, k = None, [], 0
    for r in rows:
        if r and r[0] == 'Kernel Name':
            k += 1
            if k == 2:
                break
            print(r[1][:160])
            continue
        if r and r[0] == 'Address':
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    iS, iN, iSt = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[iN]) for r in data)
    print(f'SASS lines {len(data)}, warp instructions {tot} = {tot * 32 / nsamp:.2f} per sample (one lane per sequence)')
    cnt, st = collections.Counter(), collections.Counter()
    for r in data:
        t = r[iS].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        cnt[op] += int(r[iN])
        st[op] += int(r[iSt])
    sts = sum(st.values())
    print('opcode      share   per sample   share of warp-stall samples')
    for op, c in cnt.most_common(20):
        print(f'{op:10s} {c / tot * 100:5.1f}%   {c * 32 / nsamp:6.2f}        {st[op] / sts * 100:5.1f}%')
    print('most-stalled instructions (stall samples, times executed, instruction | dominant reasons):')
    for r in sorted(data, key=lambda r: -int(r[iSt]))[:12]:
        why = ' '.join(f'{hdr[i][6:]}={r[i]}' for i in stall_cols if r[i] not in ('0', '') and int(r[i]) > 0.15 * int(r[iSt]))
        print(f'  {int(r[iSt]):6d} ({int(r[iSt]) / sts * 100:4.1f}%) {r[iN]:>8s}  {r[iS].strip()[:70]:70s} | {why}')


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 65536.0 * 4096.0)
