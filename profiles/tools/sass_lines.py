#!/usr/bin/env python
"""
Static SASS footprint per source line of one kernel (no GPU needed):

    cuobjdump -xelf all LIB.so && nvdisasm -g d3d_api.sm_100a.cubin > all.sass
    python profiles/tools/sass_lines.py all.sass KERNEL_SUBSTRING [--top N] [--ranges a-b,c-d]

Counts the instructions nvdisasm attributes to every (file, line) of the kernel's .text section
(inlined callees included), and optionally the totals over ranges of lines of the kernel's own file:
the tool behind the "instructions per role" tables of profiles/r02_notes.md (16 bytes each).
"""
import collections
import re
import sys


def main():
    path, key = sys.argv[1], sys.argv[2]
    top, ranges = 25, []
    args = sys.argv[3:]
    while args:
        a = args.pop(0)
        if a == '--top':
            top = int(args.pop(0))
        elif a == '--ranges':
            ranges = [tuple(int(v) for v in r.split('-')) for r in args.pop(0).split(',')]
    inside = False
    cur = ('?', 0)
    counts = collections.Counter()
    opc = collections.defaultdict(collections.Counter)
    n = 0
    with open(path) as f:
        for line in f:
            if line.startswith('\t.section'):
                inside = ('.text.' in line) and (key in line)
                continue
            if not inside:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
            if m:
                cur = (m.group(1).split('/')[-1], int(m.group(2)))
                continue
            m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
            if m:
                counts[cur] += 1
                opc[cur][m.group(2).split('.')[0]] += 1
                n += 1
    print('%d instructions = %.1f KB' % (n, n * 16 / 1024.0))
    for (fn, ln), c in counts.most_common(top):
        ops = ' '.join('%s:%d' % kv for kv in opc[(fn, ln)].most_common(4))
        print('%6d  %s:%d   %s' % (c, fn, ln, ops))
    if ranges:
        files = collections.Counter()
        for (fn, ln), c in counts.items():
            files[fn] += c
        main_file = files.most_common(1)[0][0]
        print('ranges of %s:' % main_file)
        for a, b in ranges:
            tot = sum(c for (fn, ln), c in counts.items() if fn == main_file and a <= ln <= b)
            print('  lines %d-%d: %d instructions' % (a, b, tot))
        for fn, c in files.most_common(8):
            print('  file %s: %d' % (fn, c))


if __name__ == '__main__':
    main()
