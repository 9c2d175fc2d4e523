#!/usr/bin/env python
"""Totals of the IFT level launches of the last bench step from an
`ncu --metrics gpu__time_duration.sum,lts__t_requests_srcunit_tex_op_atom_dot_alu.sum,...cas.sum,dram__bytes_*,smsp__inst_executed.sum
 -k regex:ift_level_kernel --csv` log.  usage: ncu_ift_levels.py <log.csv> <levels per run> <out.txt>"""
import collections
import csv
import re
import sys

M = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
     'lts__t_requests_srcunit_tex_op_atom_dot_alu.sum', 'lts__t_requests_srcunit_tex_op_atom_dot_cas.sum',
     'smsp__inst_executed.sum']


def main(path, levels, out):
    lines = [l for l in open(path) if not l.startswith('==')]
    by = collections.OrderedDict()
    for r in csv.DictReader(lines):
        by.setdefault(r['ID'], {'k': r['Kernel Name']})[r['Metric Name']] = (
            float(r['Metric Value'].replace(',', '')), r['Metric Unit'])
    last = list(by)[-int(levels):]

    def val(i, m):
        v, u = by[i][m]
        if m == 'gpu__time_duration.sum':
            v = v / 1e6 if u == 'ns' else v / 1e3 if u == 'us' else v
        if 'bytes' in m:
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
        return v
    agg = collections.OrderedDict()
    for i in last:
        g = re.search(r'ift_level_kernel<(\d+)>', by[i]['k'])
        a = agg.setdefault('ift_level_kernel<%s>' % (g.group(1) if g else '?'), [0] + [0.0] * len(M))
        a[0] += 1
        for j, m in enumerate(M):
            a[j + 1] += val(i, m)
    tot = [sum(a[j] for a in agg.values()) for j in range(len(M) + 1)]
    with open(out, 'w') as f:
        f.write('# ncu --metrics ' + ','.join(M) + '\n#   --clock-control none -k regex:ift_level_kernel, bench.py cfg3 '
                '(10M facade, k=32): the %s level launches of the last step\n# (serialised under ncu: programmatic '
                'dependent launch does not overlap here; compare shares, not absolutes)\n' % levels)
        f.write('%-24s %8s %9s %12s %12s %14s %14s %14s\n' % ('kernel', 'launches', 'ms', 'dram_rd_MB', 'dram_wr_MB',
                                                          'atomMin_req', 'atomCAS_req', 'warp_inst'))
        for k, a in sorted(agg.items(), key=lambda kv: -int(re.search(r'<(\d+|\?)>', kv[0]).group(1).replace('?', '0'))) + [('TOTAL', tot)]:
            f.write('%-24s %8d %9.3f %12.1f %12.1f %14d %14d %14d\n' % (k, a[0], a[1], a[2] / 1e6, a[3] / 1e6, a[4], a[5], a[6]))
        f.write('\nL2 atomic requests per run: %d (atomicMin offers %d + atomicCAS hooks %d)\n' % (tot[4] + tot[5], tot[4], tot[5]))
        f.write('DRAM traffic %.2f GB per run\n' % ((tot[2] + tot[3]) / 1e9))
    print(open(out).read())


if __name__ == '__main__':
    main(*sys.argv[1:])
