"""SASS instruction mix of the dominant generated kernels (runs WITHOUT a GPU: programs are compiled for sm_100a by
NVRTC into a scratch cache directory and disassembled with cuobjdump).  Writes profiles/r02_sass_mix.md.
Usage: python scripts/sass_mix.py"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(which, cache):
    """Runs in a child process (the cache directory is fixed at first use of the library)."""
    os.environ['NGB200_CACHE_DIR'] = cache
    import torch
    from numga_b200 import B200Context
    if which.startswith('chain'):
        from numga_b200.examples.physics import ChainStep, setup_bodies
        ctx = B200Context('x+y+z+w0', dtype=torch.float64 if which.endswith('64') else torch.float32, device='cpu', compile_only=True)
        bodies, sets = setup_bodies(ctx, n_bodies=12)
        for f in list(os.listdir(cache)):
            os.remove(os.path.join(cache, f))
        ChainStep(bodies, sets)
        return
    spec = {
        'sandwich': ((3, 0, 1), torch.float32, lambda c: (lambda r, x: r >> x, [(c.subspace.even_grade(), 1), (c.subspace.antivector(), 8)])),
        'gp': ((3, 0, 1), torch.float32, lambda c: (lambda a, b: a * b, [(c.subspace.even_grade(), 8), (c.subspace.even_grade(), 8)])),
        'roundtrip': ((3, 0, 1), torch.float32, lambda c: (lambda b: b.exp().normalized().motor_log(), [(c.subspace.bivector(), 8)])),
        'roundtrip64': ((3, 0, 1), torch.float64, lambda c: (lambda b: b.exp().normalized().motor_log(), [(c.subspace.bivector(), 8)])),
        'cga': ((4, 1, 0), torch.float32, lambda c: (lambda a, b: a * b, [(c.subspace.full(), 8), (c.subspace.full(), 8)])),
    }[which]
    ctx = B200Context(spec[0], dtype=spec[1], device='cpu', compile_only=True)
    fn, args = spec[2](ctx)
    mvs = [ctx.multivector(values=torch.zeros((n, len(s)) if n > 1 else (len(s),), dtype=spec[1]), subspace=s) for s, n in args]
    ctx.jit(fn)(*mvs)


def mix(cubin, kernel):
    out = subprocess.run(['cuobjdump', '-sass', '-fun', kernel, cubin], capture_output=True, text=True).stdout
    ops = collections.Counter()
    for line in out.splitlines():
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            ops[m.group(2)] += 1
    res = subprocess.run(['cuobjdump', '-res-usage', cubin], capture_output=True, text=True).stdout
    regs = re.search(r'Function %s:\s*\n\s*REG:(\d+)' % kernel, res)
    return ops, int(regs.group(1)) if regs else None


def summarise(ops):
    total = sum(ops.values())
    g = lambda *pre: sum(v for k, v in ops.items() if any(k.startswith(p) for p in pre))
    keys = [('total', total), ('FFMA', g('FFMA.', 'FFMA ') + ops.get('FFMA', 0)), ('FFMA2', g('FFMA2')), ('FMUL/FADD', g('FMUL', 'FADD')),
            ('FMUL2/FADD2', g('FMUL2', 'FADD2')), ('DFMA/DMUL/DADD', g('DFMA', 'DMUL', 'DADD')), ('MUFU', g('MUFU')),
            ('LDG.E.128', g('LDG.E.128')), ('LDG other', g('LDG') - g('LDG.E.128')), ('STG.E.128', g('STG.E.128')),
            ('STG other', g('STG') - g('STG.E.128')), ('LDS', g('LDS')), ('STS', g('STS')), ('BAR', g('BAR')),
            ('LDL/STL (spill)', g('LDL', 'STL')), ('UTMALDG/UTCMMA (TMA / tcgen05)', g('UTMA', 'UTC'))]
    return keys


if __name__ == '__main__':
    if len(sys.argv) == 3:
        build(sys.argv[1], sys.argv[2])
        sys.exit(0)
    rows = []
    for which, kernel, label in (('sandwich', 'ngb_vec', 'config 2: R >> p, broadcast motor, fp32'),
                                 ('gp', 'ngb_vec', 'config 1: motor * motor, fp32'),
                                 ('roundtrip', 'ngb_vec', 'config 3: exp -> normalized -> motor_log, fp32'),
                                 ('roundtrip64', 'ngb_vec', 'config 3: same, fp64'),
                                 ('cga', 'ngb_vec', 'config 4: CGA full x full, fp32'),
                                 ('chain', 'ngb_pipeline', 'config 5: ChainStep pipeline kernel, fp32'),
                                 ('chain64', 'ngb_pipeline', 'config 5: ChainStep pipeline kernel, fp64')):
        cache = tempfile.mkdtemp(prefix='ngb_sass_')
        subprocess.run([sys.executable, os.path.abspath(__file__), which, cache], check=True)
        cubins = [os.path.join(cache, f) for f in os.listdir(cache) if f.endswith('.cubin')]
        best = None
        for c in cubins:
            ops, regs = mix(c, kernel)
            if ops and (best is None or sum(ops.values()) > sum(best[0].values())):
                best = (ops, regs)
        rows.append((label, kernel, best[1], summarise(best[0])))
        print(label, best[1], dict(summarise(best[0])), flush=True)
    with open(os.path.join(ROOT, 'profiles', 'r02_sass_mix.md'), 'w') as f:
        f.write('# SASS instruction mix of the dominant generated kernels (sm_100a, `cuobjdump -sass`, static counts)\n\n'
                'Produced by `python scripts/sass_mix.py` (no GPU needed).  Counts are static instructions of the named entry\n'
                'point; the loop bodies of `ngb_vec` run once per 4 (fp32, `LDG.E.128`) / 2 (fp64) / 1 items, `ngb_pipeline` once per tile.\n\n')
        heads = [k for k, _ in rows[0][3]]
        f.write('| kernel | entry | regs | ' + ' | '.join(heads) + ' |\n|---|---|---|' + '---|' * len(heads) + '\n')
        for label, kernel, regs, keys in rows:
            f.write(f'| {label} | `{kernel}` | {regs} | ' + ' | '.join(str(v) for _, v in keys) + ' |\n')
        f.write('\nReading: the HBM-bound kernels move data only with 128-bit `LDG.E.128` / `STG.E.128` (one per component row per 4 items);\n'
                'no kernel contains `UTMALDG` / `UTCMMA`: nothing is reused between threads, so there is no tile for TMA to stage and no\n'
                'contraction for tcgen05 (DESIGN.md "Tensor cores").  `FFMA2` appears only in the small HBM-bound programs (packed pairs);\n'
                'the compute-bound chains are scalar `FFMA` (same 128 lanes/clk/SM issue rate on B200, fewer registers).\n')
