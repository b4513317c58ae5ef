"""Warp-stall samples per source line from an .ncu-rep captured with --import-source on (needs -lineinfo builds).

    python profiles/ncu_hotlines.py gpurun_out/prof_x.ncu-rep [top_n [kernel-regex]] > profiles/rN_ncu_x_hotlines.txt
"""
import collections
import csv
import io
import subprocess
import sys


def main(path, top=30, kernel=None):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass']
                         + (['--kernel-name-base', 'demangled', '--kernel-name', 'regex:' + kernel] if kernel else []),
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = [r for r in rows if r and r[0] == 'Line No'][0]
    col = {}
    for i, n in enumerate(hdr):
        col.setdefault(n, i)
    stalls = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]

    def num(x):
        try:
            return float(x.replace(',', ''))
        except ValueError:
            return 0.0
    cur, cur_file, func = None, None, None
    agg = collections.defaultdict(lambda: [0.0, collections.Counter()])
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
        elif r[0] == 'Function Name':
            func = r[1]
        elif r[0].strip().isdigit():
            cur = (cur_file, int(r[0]), ','.join(r[1:4]).strip()[:95])
        elif len(r) == len(hdr) and r[0] == '' and r[2].startswith('0x') and cur:
            a = agg[cur]
            a[0] += num(r[col['# Samples']])
            for st in stalls:
                a[1][st] += num(r[col[st]])
    total = sum(a[0] for a in agg.values()) or 1.0
    print('ncu source view of %s: warp-stall samples per source line (top %d of %d samples), three largest stall reasons'
          % (func, top, total))
    for (f, ln, txt), (s, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print('%5.1f%%  %s:%-4d %-95s %s' % (100 * s / total, f, ln, txt,
                                            ' '.join('%s:%d' % (k[6:], v) for k, v in st.most_common(3))))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30, sys.argv[3] if len(sys.argv) > 3 else None)
