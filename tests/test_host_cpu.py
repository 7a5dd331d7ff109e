"""CPU-side tests: the C-ABI library loads and exports every declared symbol, the drop-in module
tree reproduces the reference state_dict layout, and host-only entry points behave."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import fastscnn_oracle as fo
import metric_oracle as mo
from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from fscnn_b200 import native
    header = open(os.path.join(ROOT, 'include', 'fscnn_b200.h')).read()
    declared = set(re.findall(r'\b(fscnn_[a-z_0-9]+)\s*\(', header))
    declared -= {'fscnn_ctx', 'fscnn_tensor', 'fscnn_tap'}
    lib = native.lib()
    for sym in sorted(declared):
        assert hasattr(lib, sym), sym
    assert set(native.EXPORTED_SYMBOLS) == declared
    assert lib.fscnn_abi_version() == 1


@pytest.mark.parametrize('nc,aux', [(19, False), (2, True)])
def test_state_dict_layout_matches_reference(nc, aux):
    from models.fast_scnn import FastSCNN
    model = FastSCNN(nc, aux=aux)
    spec = fo.state_dict_spec(nc, aux)   # verified against the reference by tests/test_oracle_golden.py fixtures
    sd = model.state_dict()
    assert list(sd.keys()) == [name for name, _, _ in spec]
    for name, shape, _ in spec:
        assert tuple(sd[name].shape) == tuple(shape), name
    assert len(sd) == (276 if aux else 268)
    # attribute paths the reference's callers introspect
    assert model.classifier.conv[1].out_channels == nc
    assert model.global_feature_extractor.ppm.conv1.conv[0].in_channels == 128
    assert callable(model.global_feature_extractor.ppm.pool)


def test_native_manifest_matches_state_dict():
    from fscnn_b200 import native
    lib = native.lib()
    for nc, aux in ((19, 0), (3, 1)):
        ctx = C.c_void_p()
        native.check(lib.fscnn_create(C.byref(ctx), nc, aux, native.PREC_FP32))
        names = {lib.fscnn_param_name(ctx, i).decode(): lib.fscnn_param_numel(ctx, i) for i in range(lib.fscnn_param_count(ctx))}
        spec = {n: int(np.prod(s)) for n, s, k in fo.state_dict_spec(nc, bool(aux)) if k != 'bn_n'}
        assert names == spec
        nbytes = C.c_size_t()
        native.check(lib.fscnn_workspace_bytes(ctx, 1, 1024, 2048, C.byref(nbytes)))
        assert 100e6 < nbytes.value < 400e6
        native.check(lib.fscnn_workspace_bytes(ctx, 1000, 1024, 2048, C.byref(nbytes)))   # capped by the micro-batch
        assert nbytes.value < 128 * 400e6                                               # default micro-batch <= 128 images
        assert lib.fscnn_workspace_bytes(ctx, 1, 2, 2, C.byref(nbytes)) < 0
        assert b'bad shape' in lib.fscnn_last_error()
        tap = native.Tap()
        native.check(lib.fscnn_tap_info(ctx, 1, 1024, 2048, b'l2d.conv', C.byref(tap)))
        assert (tap.h, tap.w, tap.c) == (511, 1023, 32)
        native.check(lib.fscnn_tap_info(ctx, 1, 360, 640, b'gfe.ppm', C.byref(tap)))
        assert (tap.h, tap.w, tap.c) == (12, 20, 128)
        assert lib.fscnn_tap_info(ctx, 1, 360, 640, b'nope', C.byref(tap)) < 0
        lib.fscnn_destroy(ctx)
    assert lib.fscnn_create(C.byref(ctx), 0, 0, 0) < 0


def test_checkpoint_variants_load():
    from models.fast_scnn import FastSCNN, get_fast_scnn
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in fo.make_state_dict(2, False, 3).items()}
    FastSCNN(2).load_state_dict(sd)
    FastSCNN(2).load_state_dict({'module.' + k: v for k, v in sd.items()})
    FastSCNN(2).load_state_dict({'state_dict': sd, 'epoch': 3})
    FastSCNN(2, num_class=7)                       # unknown kwargs are swallowed like the reference
    assert get_fast_scnn('tusimple').num_classes == 2 and get_fast_scnn('citys').num_classes == 19
    with pytest.raises(TypeError):
        get_fast_scnn('citys', num_classes=3)      # positional clash, same as the reference (Appendix E)
    with pytest.raises(KeyError):
        get_fast_scnn('bdd100k', pretrained=True)  # no acronym, same as the reference


def test_cpu_or_train_mode_raises():
    from models.fast_scnn import FastSCNN
    m = FastSCNN(2).eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError):          # training mode runs the CUDA training operators: no CPU path either
        m.train()(torch.zeros(1, 3, 64, 64))


def test_conf_to_totals_host_function():
    from fscnn_b200 import native
    rng = np.random.RandomState(0)
    for nc in (2, 19):
        pred = rng.randint(-1, nc + 2, size=(3, 40, 50))
        label = rng.randint(-2, nc + 2, size=(3, 40, 50))
        conf = mo.confusion_counts(pred, label, nc)
        inter, union = np.zeros(nc, np.int64), np.zeros(nc, np.int64)
        correct, labeled = C.c_longlong(), C.c_longlong()
        ll = C.POINTER(C.c_longlong)
        native.check(native.lib().fscnn_conf_to_totals(conf.ctypes.data_as(ll), nc, inter.ctypes.data_as(ll),
                                                       union.ctypes.data_as(ll), C.byref(correct), C.byref(labeled)))
        o = mo.SegmentationMetricOracle(nc)
        o.update(pred, label)
        assert np.array_equal(inter, o.total_inter) and np.array_equal(union, o.total_union)
        assert (correct.value, labeled.value) == (o.total_correct, o.total_label)


def test_palette_output_matches_reference_tables():
    """get_color_pallete keeps the reference contract (PIL 'P' image + dataset palette, utils/visualize.py:7-36)."""
    from utils.visualize import get_color_pallete, palette_for
    mask = np.arange(19, dtype=np.int64).reshape(1, 19).repeat(3, 0)
    img = get_color_pallete(mask, 'citys')
    assert img.mode == 'P' and img.size == (19, 3)
    rgb = np.asarray(img.convert('RGB'))
    assert tuple(rgb[0, 0]) == (128, 64, 128) and tuple(rgb[0, 13]) == (0, 0, 142) and tuple(rgb[0, 18]) == (119, 11, 32)
    voc = palette_for('tusimple')
    assert tuple(voc[0]) == (0, 0, 0) and tuple(voc[1]) == (128, 0, 0) and tuple(voc[2]) == (0, 128, 0) and tuple(voc[15]) == (192, 128, 128)
    assert tuple(np.asarray(get_color_pallete(np.array([[1, 2]]), 'tusimple').convert('RGB'))[0, 1]) == (0, 128, 0)


def test_palettes_equal_the_reference_tables_entry_for_entry():
    """The complete colour tables and a get_color_pallete round trip against vectors produced by the unmodified reference
    (oracle/gen_golden_visual.py)."""
    from conftest import GOLDEN
    from utils.visualize import get_color_pallete, palette_for
    g = np.load(os.path.join(GOLDEN, 'palettes.npz'))
    assert np.array_equal(palette_for('citys')[:19], g['citys'])
    assert np.array_equal(palette_for('tusimple'), g['voc']) and np.array_equal(palette_for('pascal_voc'), g['voc'])
    assert np.array_equal(np.asarray(get_color_pallete(g['cls_map'].copy(), 'citys').convert('RGB')), g['rgb_citys'])
    assert np.array_equal(np.asarray(get_color_pallete(g['cls_map'].copy(), 'tusimple').convert('RGB')), g['rgb_voc'])


def test_create_rejects_unsupported_class_counts_up_front():
    from fscnn_b200 import native
    lib = native.lib()
    ctx = C.c_void_p()
    assert lib.fscnn_create(C.byref(ctx), 129, 0, native.PREC_BF16) < 0 and b'bf16' in lib.fscnn_last_error()
    assert lib.fscnn_create(C.byref(ctx), 241, 0, native.PREC_FP32) < 0
    native.check(lib.fscnn_create(C.byref(ctx), 128, 0, native.PREC_BF16))
    lib.fscnn_destroy(ctx)
    native.check(lib.fscnn_create(C.byref(ctx), 240, 0, native.PREC_FP32))
    lib.fscnn_destroy(ctx)


def test_kernel_generation_options_and_e2e_argument_checks():
    """No GPU needed: the option keys of the bf16 kernel generations exist, unknown keys are refused, and the camera-frame
    entry points validate their arguments before touching the device."""
    from fscnn_b200 import native
    lib = native.lib()
    ctx = C.c_void_p()
    native.check(lib.fscnn_create(C.byref(ctx), 19, 0, native.PREC_BF16))
    try:
        for key in (b'front_transposed', b'fuse_front'):
            for v in (0, 1):
                assert lib.fscnn_set_option(ctx, key, v) == 0, key
        assert lib.fscnn_set_option(ctx, b's1_transposed', 0) < 0        # the superseded kernel generation is gone
        assert lib.fscnn_set_option(ctx, b'no_such_option', 1) < 0
        assert b'no_such_option' in lib.fscnn_last_error()
        nbytes = C.c_size_t()
        native.check(lib.fscnn_workspace_bytes(ctx, 111, 1024, 2048, C.byref(nbytes)))   # one wave-exact micro-batch, bf16
        assert 111 * 60e6 < nbytes.value < 111 * 120e6
    finally:
        lib.fscnn_destroy(ctx)
    assert lib.fscnn_e2e_preprocess(None, native.U8, 1, 8, 8, 16, None, None, None, None) < 0
    dummy = C.c_void_p(16)
    assert lib.fscnn_e2e_preprocess(dummy, native.I64, 1, 8, 8, 16, None, None, dummy, None) < 0          # frames are uint8 / float32
    mean = (C.c_float * 3)(0.5, 0.5, 0.5)
    assert lib.fscnn_e2e_preprocess(dummy, native.U8, 1, 8, 8, 16, mean, None, dummy, None) < 0            # mean without std
    assert lib.fscnn_e2e_postprocess(dummy, 33, 36, 1, 4, 4, 32, 32, 8, 8, 1, dummy, None) < 0             # more than 32 classes
    assert lib.fscnn_e2e_postprocess(dummy, 4, 2, 1, 4, 4, 32, 32, 8, 8, 1, dummy, None) < 0               # padded < classes


def test_loss_module_mirrors_the_reference_interface():
    """utils/loss.py: the reference's class names, constructor defaults (utils/loss.py:15, :45, :74, :104, :128, :186) and its
    call-time failures that need no device."""
    import inspect
    import pytest
    import torch
    from utils import loss as L

    def defaults(cls):
        return {k: v.default for k, v in inspect.signature(cls.__init__).parameters.items() if v.default is not inspect.Parameter.empty}

    assert L.__all__ == ['MixSoftmaxCrossEntropyLoss', 'MixSoftmaxCrossEntropyOHEMLoss', 'DiceLoss', 'MixDiceLoss']
    assert defaults(L.DiceLoss) == {'smooth': 1e-6}
    assert defaults(L.MixDiceLoss) == {'aux': True, 'aux_weight': 0.4, 'smooth': 1e-6}
    assert defaults(L.FocalDiceLoss) == {'alpha': 0.5, 'gamma': 2.0, 'dice_weight': 0.5, 'smooth': 1e-6}
    assert defaults(L.MixSoftmaxCrossEntropyLoss) == {'aux': True, 'aux_weight': 0.2, 'ignore_label': -1}
    assert defaults(L.SoftmaxCrossEntropyOHEMLoss) == {'ignore_label': -1, 'thresh': 0.7, 'min_kept': 256, 'use_weight': True}
    assert defaults(L.MixSoftmaxCrossEntropyOHEMLoss) == {'aux': False, 'aux_weight': 0.2, 'ignore_index': -1}
    assert isinstance(L.MixSoftmaxCrossEntropyLoss(), torch.nn.CrossEntropyLoss) and L.MixSoftmaxCrossEntropyLoss().ignore_index == -1
    assert issubclass(L.MixSoftmaxCrossEntropyOHEMLoss, L.SoftmaxCrossEntropyOHEMLoss)
    w = L.SoftmaxCrossEntropyOHEMLoss().weight
    assert w.shape == (19,) and abs(float(w[0]) - 0.8373) < 1e-6 and abs(float(w[18]) - 1.0507) < 1e-6
    assert L.SoftmaxCrossEntropyOHEMLoss(use_weight=False).weight is None
    pred, target = torch.zeros(1, 2, 8, 8), torch.zeros(1, 8, 8, dtype=torch.int64)
    with pytest.raises(AttributeError):      # loss.py:82: pred.dim() on the tuple train.py hands over
        L.FocalDiceLoss()((pred, pred), target)
    with pytest.raises(TypeError):           # loss.py:124: nn.CrossEntropyLoss.forward(pred0, pred1, target)
        L.MixSoftmaxCrossEntropyLoss(aux=False)((pred, pred), target)
    with pytest.raises(RuntimeError):        # CPU tensors: there is no CPU path
        L.DiceLoss()(pred, target)


def test_lr_scheduler_is_bit_identical_to_the_reference():
    """utils/lr_scheduler.LRScheduler against learning rates the reference class produced (tests/golden/lr_scheduler_cases.json, written
    by running both classes side by side in the build container): every mode, offsets, epoch-based steps; and Trainer's poly_lr agrees."""
    import json
    import os
    from conftest import GOLDEN
    from fscnn_b200 import poly_lr
    from utils.lr_scheduler import LRScheduler
    for case in json.load(open(os.path.join(GOLDEN, 'lr_scheduler_cases.json'))):
        sched = LRScheduler(**case['kwargs'])
        assert [sched(i) for i in case['iters']] == case['lr'], case['kwargs']
    sched = LRScheduler(mode='poly', base_lr=0.01, nepochs=60, iters_per_epoch=176, power=0.9)
    for it in (0, 1, 500, 60 * 176 - 2):
        assert abs(sched(it) - poly_lr(0.01, it, 60, 176)) <= 1e-15


def test_criterion_argument_checks_need_no_device():
    """train_ops.criterion / adamw_step refuse bad arguments before any kernel could see them."""
    import pytest
    import torch
    from fscnn_b200 import train_ops
    logits, target = torch.zeros(1, 2, 8, 8), torch.zeros(1, 8, 8, dtype=torch.int64)
    with pytest.raises(ValueError):
        train_ops.criterion(logits, target, 'hinge')
    with pytest.raises(RuntimeError):                      # CPU tensors: no CPU path
        train_ops.criterion(logits, target, 'ce')
    with pytest.raises(ValueError):
        train_ops.adamw_step(torch.zeros(4), torch.zeros(4), torch.zeros(4), torch.zeros(4), 1e-3, 1)
    assert train_ops.CRITERIA == {'ce': 0, 'dice': 1, 'focal_dice': 2}
