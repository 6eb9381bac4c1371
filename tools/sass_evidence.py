"""Per-kernel counts of the Blackwell-only mnemonics in libdgb200.so (cuobjdump -sass) -> profiles/r2_sass_evidence.md."""
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(REPO, 'deepgemm_b200', 'lib', 'libdgb200.so')
sass = subprocess.run(['/usr/local/cuda/bin/cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
kernels = sass.split('Function : ')[1:]
cols = ['UTCQMMA', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'STSM', 'UTCCP', 'UTCBAR', 'LDTM', 'UBLKCP', 'R2UR.BROADCAST', 'HMMA']
demangle = subprocess.run(['c++filt'], input='\n'.join(k.split('\n', 1)[0].strip() for k in kernels), capture_output=True, text=True).stdout.splitlines()
rows = []
for k, name in zip(kernels, demangle):
    body = k.split('\n', 1)[1]
    short = re.sub(r'\(.*', '', name).replace('dgb200::', '').replace('void ', '')
    rows.append((short, [len(re.findall(r'\b' + re.escape(c) + r'\b', body)) for c in cols]))
tot = [sum(r[1][i] for r in rows) for i in range(len(cols))]
out = ['# SASS evidence, round 2 (cuobjdump -sass deepgemm_b200/lib/libdgb200.so, sm_100a)', '',
       f'{len(rows)} kernels in the library, {sum(1 for r in rows if "fp8_gemm_kernel" in r[0])} of them instances of `fp8_gemm_kernel`. '
       'Template arguments: <gemm type, cluster, out type, accumulate, X MN-major, W MN-major, workspace split-K, cluster split-K slices, TMA store, transposed output>.', '',
       '| mnemonic | PTX it comes from | total in the library |', '|---|---|---|']
what = {'UTCQMMA': 'tcgen05.mma kind::mxf8f6f4.block_scale', 'UTMALDG': 'cp.async.bulk.tensor (TMA load)', 'UTMASTG': 'cp.async.bulk.tensor store (TMA store epilogue)',
        'UTMAPF': 'cp.async.bulk.prefetch.tensor (L2 prefetch at kernel entry)', 'STSM': 'stmatrix.x4.trans (store epilogue)', 'UTCCP': 'tcgen05.cp (scale factors -> TMEM)',
        'UTCBAR': 'tcgen05.commit', 'LDTM': 'tcgen05.ld', 'UBLKCP': 'cp.async.bulk shared::cluster (split-K exchange)', 'R2UR.BROADCAST': 'waterfall around a TMA operand (must be 0)',
        'HMMA': 'legacy mma.sync (must be 0)'}
for c, t in zip(cols, tot):
    out.append(f'| `{c}` | {what[c]} | {t} |')
out += ['', '| kernel | ' + ' | '.join(cols[:9]) + ' |', '|---|' + '---|' * 9]
pick = [r for r in rows if 'fp8_gemm_kernel' in r[0]]
seen = set()
for name, cnt in pick:
    key = tuple(cnt)
    tag = re.sub(r'fp8_gemm_kernel', '', name)
    if any(s in tag for s in ('<0, 2, __nv_bfloat16, false, false, false, false, 0, false, false>', '<0, 2, __nv_bfloat16, false, false, false, false, 0, true, false>',
                              '<0, 4, __nv_bfloat16, false, false, false, false, 4, false, false>', '<0, 4, __nv_bfloat16, false, false, false, false, 2, false, false>',
                              '<0, 1, __nv_bfloat16, false, false, false, false, 0, false, true>', '<1, 2, __nv_bfloat16, false, false, false, false, 0, false, false>',
                              '<2, 2,', '<3, 2, __nv_bfloat16, false, false, false', '<4, 2,', '<6, 2, __nv_bfloat16, false, false, false', '<0, 2, float, true, true, true, true')):
        out.append(f'| `{tag[:90]}` | ' + ' | '.join(str(x) for x in cnt[:9]) + ' |')
others = [r for r in rows if 'fp8_gemm_kernel' not in r[0]]
out += ['', 'Other kernels: ' + ', '.join(f'`{re.sub("<.*", "", r[0])}`' for r in others) + '.']
path = os.path.join(REPO, 'profiles', 'r2_sass_evidence.md')
open(path, 'w').write('\n'.join(out) + '\n')
print('\n'.join(out[:30]))
