"""CGA (4,1,0) full x full geometric product: which source shape does ptxas turn into the fastest kernel?
Emits scripts/micro/cga_variants.cu = hand-shaped prototypes + the generator's own sources (dumped with
NGB200_CACHE_SOURCES=1 into the directories given on the command line as name=dir), all timed in one process in
interleaved rounds (median reported), so that clock drift under sustained load does not favour a position in the list."""
import glob, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from numga_b200.algebra import Algebra

A = Algebra.from_pqr(4, 1, 0)
F = A.subspace.full()
op = A.operator.product(F, F)
rows = {}
for (i, j, o), c in zip(op.index.tolist(), op.coeff.tolist()):
    rows.setdefault(o, []).append((i, j, c))


def neg(c): return '-' if c < 0 else ''


def body_outmajor(sfx=''):
    s = ''
    for o in range(32):
        terms = sorted(rows[o], key=lambda t: t[1])
        i, j, c = terms[0]
        s += f'    float y{o}{sfx} = {neg(c)}a{sfx}[{i}] * b{sfx}[{j}];\n'
        for i, j, c in terms[1:]:
            s += f'    y{o}{sfx} = fmaf({neg(c)}a{sfx}[{i}], b{sfx}[{j}], y{o}{sfx});\n'
    return s


def body_bmajor(sfxs=('',), load_b=None):
    s = ''
    for j in range(32):
        if load_b: s += load_b(j)
        for o in range(32):
            (i, _, c), = [t for t in rows[o] if t[1] == j]
            for sfx in sfxs:
                if j == 0: s += f'    float y{o}{sfx} = {neg(c)}a{sfx}[{i}] * b{sfx}[{j}];\n'
                else: s += f'    y{o}{sfx} = fmaf({neg(c)}a{sfx}[{i}], b{sfx}[{j}], y{o}{sfx});\n'
    return s


HEAD = '__global__ void __launch_bounds__(%d) %s(const float* __restrict__ pa, const float* __restrict__ pb, float* __restrict__ out, long long n, long long pitch) {\n'
LOOP1 = '  const long long stride = (long long)gridDim.x * blockDim.x;\n  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {\n'
kernels = []   # (name, tpb, items per thread, source)

k = HEAD % (256, 'p_abab') + LOOP1 + '    float a[32], b[32];\n    #pragma unroll\n    for (int c = 0; c < 32; ++c) { a[c] = pa[c * pitch + i]; b[c] = pb[c * pitch + i]; }\n'
k += body_outmajor() + ''.join(f'    out[{o}LL * pitch + i] = y{o};\n' for o in range(32)) + '  }\n}\n'
kernels.append(('p_abab', 256, 1, k))

LD_AABB = '    float a[32], b[32];\n    #pragma unroll\n    for (int c = 0; c < 32; ++c) a[c] = pa[c * pitch + i];\n    #pragma unroll\n    for (int c = 0; c < 32; ++c) b[c] = pb[c * pitch + i];\n'
k = HEAD % (256, 'p_aabb') + LOOP1 + LD_AABB + body_outmajor() + ''.join(f'    out[{o}LL * pitch + i] = y{o};\n' for o in range(32)) + '  }\n}\n'
kernels.append(('p_aabb', 256, 1, k))

k = HEAD % (256, 'p_bmajor') + LOOP1 + LD_AABB + body_bmajor() + ''.join(f'    out[{o}LL * pitch + i] = y{o};\n' for o in range(32)) + '  }\n}\n'
kernels.append(('p_bmajor', 256, 1, k))

k = (HEAD % (256, 'p_lb2')).replace('__launch_bounds__(256)', '__launch_bounds__(256, 2)') + LOOP1 + LD_AABB + body_outmajor() + ''.join(f'    out[{o}LL * pitch + i] = y{o};\n' for o in range(32)) + '  }\n}\n'
kernels.append(('p_lb2', 256, 1, k))
k = HEAD % (256, 'p_norestrict') + LOOP1 + LD_AABB + body_outmajor() + ''.join(f'    out[{o}LL * pitch + i] = y{o};\n' for o in range(32)) + '  }\n}\n'
kernels.append(('p_norestrict', 256, 1, k.replace('const float* __restrict__', 'const float*').replace('float* __restrict__', 'float*')))

# two items per thread through 64-bit loads
LOOP2 = '  const long long stride = (long long)gridDim.x * blockDim.x;\n  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n / 2; v += stride) {\n'
LD2 = ('    float a0[32], a1[32], b0[32], b1[32];\n    #pragma unroll\n    for (int c = 0; c < 32; ++c) { const float2 t = reinterpret_cast<const float2*>(pa + c * pitch)[v]; a0[c] = t.x; a1[c] = t.y; }\n'
       '    #pragma unroll\n    for (int c = 0; c < 32; ++c) { const float2 t = reinterpret_cast<const float2*>(pb + c * pitch)[v]; b0[c] = t.x; b1[c] = t.y; }\n')
ST2 = ''.join(f'    reinterpret_cast<float2*>(out + {o}LL * pitch)[v] = make_float2(y{o}0, y{o}1);\n' for o in range(32))
for tpb in (128, 256):
    k = HEAD % (tpb, f'p_aabb_v2_{tpb}') + LOOP2 + LD2 + body_outmajor('0') + body_outmajor('1') + ST2 + '  }\n}\n'
    kernels.append((f'p_aabb_v2_{tpb}', tpb, 2, k))
k = HEAD % (128, 'p_bmajor_v2') + LOOP2 + LD2 + body_bmajor(('0', '1')) + ST2 + '  }\n}\n'
kernels.append(('p_bmajor_v2', 128, 2, k))

src = '#include <cstdio>\n#include <vector>\n#include <algorithm>\n#include <functional>\n#include <string>\n#include <cuda_runtime.h>\n'
ONLY = os.environ.get('ONLY')
if ONLY: kernels = [k for k in kernels if k[0] in ONLY.split(',')]
src += ''.join(k for *_, k in kernels)
gens = []
for arg in sys.argv[1:]:
    name, d = arg.split('=')
    g = open(glob.glob(os.path.join(d, '*.cu'))[0]).read().replace('extern "C" __global__', '__global__')
    src += f'namespace {name} {{\n{g}\nstatic const int TPB = NGB_TPB; static const int V = NGB_V;\n}}\n#undef NGB_V\n#undef NGB_TPB\n#undef NGB_BOUNDS\n#undef NGB_LD\n#undef NGB_ST\n'
    gens.append((name, 'ngb_vec' if 'ngb_vec(' in g else 'ngb_scalar'))
src += '''
struct Leg { std::string name; std::function<void()> launch; std::vector<float> ms; int regs; int ctas; };
int main() {
  long long n = 1LL << 26;
  float *a, *b, *o;
  cudaMalloc(&a, 32 * n * 4); cudaMalloc(&b, 32 * n * 4); cudaMalloc(&o, 32 * n * 4);
  float* h = (float*)malloc(32 * 4096 * 4);
  for (int i = 0; i < 32 * 4096; ++i) h[i] = (float)((i * 2654435761u) % 1000) / 500.f - 1.f;
  for (long long off = 0; off < 32 * n; off += 32 * 4096) { cudaMemcpy(a + off, h, 32 * 4096 * 4, cudaMemcpyHostToDevice); cudaMemcpy(b + off, h + 7, 32 * 4096 * 4 - 28, cudaMemcpyHostToDevice); }
  std::vector<Leg> legs;
  std::vector<double> sums;
'''
for name, tpb, items, _ in kernels:
    src += f'''  {{ int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, {name}, {tpb}, 0); cudaFuncAttributes at; cudaFuncGetAttributes(&at, {name});
    int grid = 148 * nb * 8; legs.push_back({{"{name}", [=]() {{ {name}<<<grid, {tpb}>>>(a, b, o, n, n); }}, {{}}, at.numRegs, nb}}); }}
'''
for name, kern in gens:
    src += f'''  {{ {name}::Params p; p.in[0] = {{a, n, 1, nullptr}}; p.in[1] = {{b, n, 1, nullptr}}; p.out[0] = {{o, n, 1, nullptr}}; p.scal[0] = 0; p.n = n;
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, {name}::{kern}, {name}::TPB, 0); cudaFuncAttributes at; cudaFuncGetAttributes(&at, {name}::{kern});
    int grid = 148 * nb * 8; legs.push_back({{"{name}", [=]() {{ {name}::{kern}<<<grid, {name}::TPB>>>(p); }}, {{}}, at.numRegs, nb}}); }}
'''
src += '''
  float* c = (float*)malloc(1 << 22);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<int> order(legs.size()); for (size_t k = 0; k < legs.size(); ++k) order[k] = (int)k;
  unsigned rng = 12345;
  for (int round = 0; round < 25; ++round) {
    if (round > 0) for (size_t k = legs.size() - 1; k > 0; --k) { rng = rng * 1664525u + 1013904223u; std::swap(order[k], order[(rng >> 8) % (k + 1)]); }   // random order: no leg always follows the same neighbour
    for (int idx : order) { auto& L = legs[idx];
      if (round == 0) { cudaMemset(o, 0, 32 * n * 4); L.launch(); double s = 0; for (int r = 0; r < 32; r += 5) { cudaMemcpy(c, o + r * n + 12345 * 4, 1 << 22, cudaMemcpyDeviceToHost); for (int i = 0; i < (1 << 20); ++i) s += c[i] * (1 + (i & 7)); } sums.push_back(s); }
      L.launch();
      cudaEventRecord(e0); for (int i = 0; i < 3; ++i) L.launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); L.ms.push_back(ms / 3);
    }
  }
  for (size_t k = 0; k < legs.size(); ++k) { auto& L = legs[k]; std::sort(L.ms.begin(), L.ms.end());
    printf("%-16s regs %3d CTAs/SM %d  median %.3f ms %5.0f GB/s  (min %.3f max %.3f)  checksum %.6e  %s\\n", L.name.c_str(), L.regs, L.ctas, L.ms[12], 384.0 * n / L.ms[12] / 1e6, L.ms.front(), L.ms.back(), sums[k], cudaGetErrorString(cudaGetLastError())); }
  return 0;
}
'''
open(os.path.join(HERE, 'cga_variants.cu'), 'w').write(src)
print('written', len(src))
