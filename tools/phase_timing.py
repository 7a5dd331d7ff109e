"""Debug utility (library built with EXTRA=-DFSCNN_PHASE_TIMING): prints the clock64 deltas between the phases of one
CTA (its second tile) of the persistent <64,64,1> bf16 bottleneck kernel while the GPU is busy with a full batch."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fast-scnn-pytorch_b200'))
import torch
from models.fast_scnn import FastSCNN
from fscnn_b200 import native

dev = torch.device('cuda', 0)
model = FastSCNN(19, precision='bf16').eval().to(dev)
x = torch.randn(int(sys.argv[1]) if len(sys.argv) > 1 else 8, 3, 1024, 2048, device=dev)
for _ in range(3):
    model.predict(x)
torch.cuda.synchronize()
lib = C.CDLL(native.lib_path())
buf = (C.c_longlong * 64)()
have_bneck = hasattr(lib, 'fscnn_debug_bneck_phases') and lib.fscnn_debug_bneck_phases(buf) == 0
t = list(buf)
if have_bneck:
    nch = 6
    names = []
    for e in range(nch):
        names += [f'chunk {e}: wait expand MMA', f'chunk {e}: expand epilogue (+barrier)', f'chunk {e}: depthwise (+proj wait, barrier)']
    names += ['wait last project MMA', 'output epilogue + store']
    for i, nme in enumerate(names):
        print(f'{nme:44s} {t[i + 1] - t[i]:8d} cycles')
    print(f'{"total":44s} {t[len(names)] - t[0]:8d} cycles')

fbuf = (C.c_longlong * 16)()
if hasattr(lib, 'fscnn_debug_front_phases') and lib.fscnn_debug_front_phases(fbuf) == 0:
    f = list(fbuf)
    print('fused front kernel, one tile of one CTA (2 CTAs share the SM):')
    fn = ['wait patch (+sync)', 'issue next prefetch', 'repack raw -> RGBX planes (+sync)', 'stem MMAs issue + wait', 'stem epilogue (+sync)',
          'depthwise (+sync)', 'pointwise MMA issue + wait', 'output epilogue']
    for i, nme in enumerate(fn):
        print(f'{nme:44s} {f[i + 1] - f[i]:8d} cycles')
    print(f'{"total":44s} {f[8] - f[0]:8d} cycles')

gbuf = (C.c_longlong * 16)()
if hasattr(lib, 'fscnn_debug_ffm_phases') and lib.fscnn_debug_ffm_phases(gbuf) == 0:
    g = list(gbuf)
    print('FFM kernel, one tile (thread 0 of one CTA):')
    gn = ['barrier: U(t-1) consumed', 'U: bilinear resize', 'barrier: U complete', 'depthwise', 'arrive + tap table', 'epilogue(t-1)']
    for i, nme in enumerate(gn):
        print(f'{nme:44s} {g[i + 1] - g[i]:8d} cycles')
    print(f'{"total":44s} {g[6] - g[0]:8d} cycles')

sbuf = (C.c_longlong * 64)()
if hasattr(lib, 'fscnn_debug_s1_phases') and lib.fscnn_debug_s1_phases(sbuf) == 0:
    s1 = list(sbuf)
    t0 = s1[0]
    print('stride-1 bottleneck pipeline, chunks 8..11 of one CTA (cycles since the first stamp):')
    print('  expand warp : loop top | expand MMA done | E buffer free | E written      depthwise warp: loop top | E ready | FMAs done | D written')
    for c in range(4):
        r = [v - t0 for v in s1[c * 8:c * 8 + 8]]
        print(f'  chunk {8 + c}: {r[0]:8d} {r[1]:8d} {r[2]:8d} {r[3]:8d}      {r[4]:8d} {r[5]:8d} {r[6]:8d} {r[7]:8d}')

tbuf = (C.c_longlong * 384)()
if hasattr(lib, 'fscnn_debug_s1t_phases') and lib.fscnn_debug_s1t_phases(tbuf) == 0 and any(v > 0 for v in tbuf):
    t = list(tbuf)
    t0 = min(v for v in t if v > 0)
    print('transposed stride-1 bottleneck, chunks 12..19 of CTA 5 (cycles since the first stamp)')
    print('  warp0: top | exp done | ldtm+cvt | fma done | D free | D written || w15 written || ctl: proj issue..end | exp issue..end | exp seen done | proj seen done | We prefetch')
    for c in range(8):
        r = [(v - t0 if v > 0 else -1) for v in t[c * 16:c * 16 + 16]]
        print(f'  chunk {12 + c}: ' + ' '.join(f'{v:7d}' for v in r[:6]) + ' || ' + f'{r[6]:7d}' + ' || ' + ' '.join(f'{v:7d}' for v in r[8:15]))
    print('  per-warp loop top (cycles since first stamp), warps 0..15, then per-warp D written')
    for c in range(8):
        print(f'  chunk {12 + c} top: ' + ' '.join(f'{(v - t0 if v > 0 else -1):6d}' for v in t[256 + c * 16:256 + c * 16 + 16]))
        print(f'  chunk {12 + c} Dwr: ' + ' '.join(f'{(v - t0 if v > 0 else -1):6d}' for v in t[128 + c * 16:128 + c * 16 + 16]))

ubuf = (C.c_longlong * 192)()
if hasattr(lib, 'fscnn_debug_s2t_phases') and lib.fscnn_debug_s2t_phases(ubuf) == 0 and any(v > 0 for v in ubuf):
    t = list(ubuf)
    t0 = min(v for v in t if v > 0)
    print('transposed stride-2 bottleneck <64,64>, units 24..35 (= chunks 6..8 x 4 sub-tiles) of CTA 5 (cycles since the first stamp)')
    print('  warp0: top | exp done | ldtm+cvt, tmfree | fma done | D free | dready | w15 top | w15 tmfree || ctl: top | tm free | committed | reloads issued | proj: D ready || epilogue end')
    for u in range(12):
        r = [(v - t0 if v > 0 else -1) for v in t[u * 16:u * 16 + 16]]
        print(f'  unit {24 + u}: ' + ' '.join(f'{v:7d}' for v in r[:8]) + ' || ' + ' '.join(f'{v:7d}' for v in r[8:13]) + ' || ' + f'{r[13]:7d}')

vbuf = (C.c_longlong * 128)()
if hasattr(lib, 'fscnn_debug_front_t_phases') and lib.fscnn_debug_front_t_phases(vbuf) == 0 and any(v > 0 for v in vbuf):
    t = list(vbuf)
    t0 = min(v for v in t if v > 0)
    print('transposed front kernel (fp32 input), tiles 8..15 of CTA 5 (cycles since the first stamp)')
    print('  stem ctl: top | planes ready | patch(t+2) issued | tm free | committed || warp0: top | stem done | ldtm+cvt | fma done | patch(t+2) here | repacked | D + planes published | epilogue(t-1) done')
    for k in range(8):
        r = [(v - t0 if v > 0 else -1) for v in t[k * 16:k * 16 + 16]]
        print(f'  tile {8 + k}: ' + ' '.join(f'{v:7d}' for v in r[:5]) + ' || ' + ' '.join(f'{v:7d}' for v in r[5:13]))

