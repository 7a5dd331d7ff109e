"""Golden vectors for the output side of the path, produced by the UNMODIFIED reference in the
build container (TEST INFRASTRUCTURE):  python oracle/gen_golden_visual.py
  palettes.npz : the reference's complete colour tables (utils/visualize.py: cityspallete, vocpallete = _getvocpallete(256)) and
                 a get_color_pallete() round trip;
  overlay.npz  : create_overlay() of demo_tusimple.py:87-104 on a seeded frame / lane mask for three alphas (the function is
                 taken from the reference file's own source; the module itself does not import here because of its script-level
                 dependencies)."""
import ast
import os
import sys

import numpy as np

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
sys.dont_write_bytecode = True
sys.path.insert(0, REF)


def main():
    from utils import visualize as rv
    citys = np.asarray(rv.cityspallete, dtype=np.uint8).reshape(-1, 3)
    voc = np.asarray(rv.vocpallete, dtype=np.uint8).reshape(-1, 3)
    rng = np.random.RandomState(0)
    cls_map = rng.randint(0, 19, size=(7, 9)).astype(np.int64)
    rgb_citys = np.asarray(rv.get_color_pallete(cls_map.copy(), 'citys').convert('RGB'))
    rgb_voc = np.asarray(rv.get_color_pallete(cls_map.copy(), 'tusimple').convert('RGB'))
    np.savez_compressed(os.path.join(OUT, 'palettes.npz'), citys=citys, voc=voc, cls_map=cls_map, rgb_citys=rgb_citys, rgb_voc=rgb_voc)
    print('palettes', citys.shape, voc.shape)

    src = open(os.path.join(REF, 'demo_tusimple.py')).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'create_overlay')
    ns = {'np': np, 'Image': __import__('PIL.Image', fromlist=['Image'])}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), 'demo_tusimple.py', 'exec'), ns)
    frame = rng.randint(0, 256, size=(24, 40, 3)).astype(np.uint8)
    lane = (rng.rand(24, 40) < 0.3)
    out = {'frame': frame, 'lane': lane}
    for alpha in (0.5, 0.3, 0.85):
        out[f'overlay_{alpha}'] = ns['create_overlay'](frame, (lane * 255).astype(np.uint8), alpha=alpha)
    np.savez_compressed(os.path.join(OUT, 'overlay.npz'), **out)
    print('overlay ok')


if __name__ == '__main__':
    main()
