"""Static issue model of a kernel from its SASS control words (cuobjdump -sass): sum of the stall counts ptxas
encoded (bits 41-44 of the second 64-bit word) = the fewest cycles ONE warp needs to issue the kernel once, beside
the instruction count and the FP64-pipe cycles (2 per FP64 instruction on a 16-lane-per-scheduler pipe).
usage: python tools/sass_stalls.py lib.so kernel_name_fragment"""
import collections
import re
import subprocess
import sys

FP64 = ('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX')


def kernel_sass(lib, frag):
    txt = subprocess.run(['cuobjdump', '-sass', lib], stdout=subprocess.PIPE, text=True).stdout
    for part in re.split(r'\n\s*Function : ', txt)[1:]:
        if frag in part.split('\n', 1)[0]:
            return part
    raise SystemExit(f'no kernel matching {frag}')


def instructions(part):
    """[(offset, opcode, text, stall, yield_, wait_mask)]"""
    out = []
    lines = part.split('\n')
    for i, line in enumerate(lines):
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+((?:@!?U?P\w+\s+)?)([A-Z0-9_.]+)(.*?);\s*/\* 0x([0-9a-f]{16}) \*/', line)
        if not m:
            continue
        m2 = re.search(r'/\* 0x([0-9a-f]{16}) \*/', lines[i + 1])
        hi = int(m2.group(1), 16)
        out.append((int(m.group(1), 16), m.group(3), m.group(2) + m.group(3) + m.group(4), (hi >> 41) & 0xf, (hi >> 45) & 1,
                    (hi >> 52) & 0x3f))
    return out


if __name__ == '__main__':
    ins = instructions(kernel_sass(sys.argv[1], sys.argv[2]))
    n = len(ins)
    stall = sum(i[3] for i in ins)
    f64 = sum(1 for i in ins if i[1].split('.')[0] in FP64)
    print(f'instructions {n}  fp64 {f64}  sum of stall counts {stall}  ({stall / n:.2f} per instruction)  fp64 pipe cycles {2 * f64}')
    h = collections.Counter(i[3] for i in ins)
    print('stall histogram', sorted(h.items()))
    hf = collections.Counter(i[3] for i in ins if i[1].split('.')[0] in FP64)
    print('  after FP64   ', sorted(hf.items()))
    w = sum(1 for i in ins if i[5])
    print(f'instructions waiting on a scoreboard: {w}')


def loops(ins):
    """innermost backward branches: (start offset, end offset, instructions, fp64, stall sum)"""
    by_off = {i[0]: n for n, i in enumerate(ins)}
    out = []
    for n, i in enumerate(ins):
        if i[1].startswith('BRA'):
            m = re.search(r'0x([0-9a-f]+)', i[2])
            if m and int(m.group(1), 16) <= i[0] and int(m.group(1), 16) in by_off:
                a = by_off[int(m.group(1), 16)]
                body = ins[a:n + 1]
                out.append((ins[a][0], i[0], len(body), sum(1 for b in body if b[1].split('.')[0] in FP64),
                            sum(b[3] for b in body), collections.Counter(b[1].split('.')[0] for b in body)))
    return out
