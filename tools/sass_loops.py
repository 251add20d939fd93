"""Offline SASS inspection: for every kernel whose (mangled) name matches the pattern, list the loops (backward
branches) with their body size and opcode mix -- used to keep the hot loops of flr_tc lean without a GPU.

    python tools/sass_loops.py afcm_b200/csrc/_obj/flr_tc.o 'flr_tc_kernelILi2ELi2E6__halfS1_Li0'
"""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    funcs = re.split(r'\n\s*Function : ', out)[1:]
    for f in funcs:
        name = f.split('\n', 1)[0].strip()
        if not re.search(pat, name):
            continue
        insts = []
        for line in f.split('\n'):
            m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
            if m:
                insts.append((int(m.group(1), 16), m.group(2).strip()))
        addr_index = {a: i for i, (a, _) in enumerate(insts)}
        print(f'== {name}: {len(insts)} instructions')
        for i, (a, s) in enumerate(insts):
            m = re.search(r'\bBRA\b.*?(0x[0-9a-f]+)', s)
            if m:
                t = int(m.group(1), 16)
                if t <= a and t in addr_index:
                    body = insts[addr_index[t]:i + 1]
                    mix = collections.Counter()
                    for _, b in body:
                        toks = b.split()
                        op = toks[1] if toks[0].startswith('@') else toks[0]
                        key = '.'.join(op.split('.')[:2]) if op.startswith(('IMAD', 'HMMA', 'HMUL2', 'HFMA2')) else op.split('.')[0]
                        mix[key] += 1
                    if mix.get('HMMA.16816', 0) or len(body) > 100:
                        print(f'  loop @{t:#x}..{a:#x}: {len(body)} instrs, ' + ', '.join(f'{k} {v}' for k, v in mix.most_common(14)))


if __name__ == '__main__':
    main()
