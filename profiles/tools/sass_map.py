#!/usr/bin/env python
"""
Address map of one kernel's SASS by source line (no GPU needed): for every 2 KB of code, the
source lines that dominate it.  Shows which roles of a warp-specialised kernel are laid out where
and how large the per-site (hot) footprint is against the instruction caches.

    nvdisasm -g d3d_api.sm_100a.cubin > all.sass
    python profiles/tools/sass_map.py all.sass KERNEL_SUBSTRING [--block 2048]
"""
import collections
import re
import sys


def main():
    path, key = sys.argv[1], sys.argv[2]
    block = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[3] == '--block' else 2048
    inside, cur, rows = False, None, []
    for line in open(path):
        if line.startswith('\t.section'):
            inside = ('.text.' in line) and (key in line)
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*?);', line)
        if m and cur:
            rows.append((int(m.group(1), 16), cur))
    blocks = collections.defaultdict(collections.Counter)
    for a, c in rows:
        blocks[a // block][c] += 1
    for b in sorted(blocks):
        top = blocks[b].most_common(3)
        print('%6.1f KB  %s' % (b * block / 1024.0, '  '.join('%s:%d(%d)' % (c[0][:14], c[1], n) for c, n in top)))


if __name__ == '__main__':
    main()
