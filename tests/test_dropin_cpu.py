"""CPU: the host mirror keeps the reference's drop-in contract: same constructor / function signatures, same
`state_dict` keys, shapes and `parameters()` order for the shipped 2023 and 2024 `net` configs, and the recipes' import lines
resolve.  The contract is the committed fixture tests/golden/api_contract.json, minted from the live reference by
oracle/make_api_contract.py; when /root/reference is present (build container) the fixture itself is re-checked against it."""
import inspect
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONTRACT = json.load(open(os.path.join(ROOT, "tests", "golden", "api_contract.json")))


def sig(fn):
    return [[p.name, str(p.kind), None if p.default is inspect.Parameter.empty else repr(p.default)]
            for p in inspect.signature(fn).parameters.values()]


def mirror():
    from desed_task_b200 import data_augm
    from desed_task_b200.nnet.CNN import CNN, GLU, ContextGating
    from desed_task_b200.nnet.CRNN import CRNN
    from desed_task_b200.nnet.RNN import BidirectionalGRU
    from desed_task_b200.utils.postprocess import ClassWiseMedianFilter
    from desed_task_b200.utils.scaler import TorchScaler
    from desed_task_b200.utils.schedulers import ExponentialWarmup
    return {"CRNN.__init__": CRNN.__init__, "CRNN.forward": CRNN.forward, "CNN.__init__": CNN.__init__,
            "GLU.__init__": GLU.__init__, "ContextGating.__init__": ContextGating.__init__,
            "BidirectionalGRU.__init__": BidirectionalGRU.__init__, "mixup": data_augm.mixup,
            "frame_shift": data_augm.frame_shift, "add_noise": data_augm.add_noise,
            "TorchScaler.__init__": TorchScaler.__init__, "ExponentialWarmup.__init__": ExponentialWarmup.__init__,
            "ClassWiseMedianFilter.__init__": ClassWiseMedianFilter.__init__}


@pytest.mark.parametrize("name", sorted(CONTRACT["signatures"]))
def test_signatures_match_the_reference(name):
    assert sig(mirror()[name]) == CONTRACT["signatures"][name], name


@pytest.mark.parametrize("tag", ["2023", "2024"])
def test_state_dict_and_parameter_order_match_the_reference(tag):
    from desed_task_b200.nnet.CRNN import CRNN
    ref = CONTRACT["nets"][tag]
    net = CRNN(**ref["config"])                     # CRNN(**config["net"]) with the shipped YAML, unknown keys included
    got = [[k, list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()]
    assert got == ref["state_dict"]
    assert [n for n, _ in net.named_parameters()] == ref["parameters"]
    assert sum(p.numel() for p in net.parameters()) == ref["n_params"]
    # deepcopy + zip(parameters) is how the recipes build and update the teacher (sed_trainer.py:62,198)
    import copy
    teacher = copy.deepcopy(net)
    assert all(a.shape == b.shape for a, b in zip(teacher.parameters(), net.parameters()))
    # a reference-shaped checkpoint loads strictly
    sd = {k: torch.zeros(shape, dtype=getattr(torch, dt.split(".")[1])) for k, shape, dt in ref["state_dict"]}
    net.load_state_dict(sd, strict=True)


@pytest.mark.parametrize("name", sorted(CONTRACT["SEDTask4"]))
def test_lightning_module_surface_matches_the_2023_recipe(name):
    """SEDTask4 hot-path methods: same positional parameters and defaults as recipes/dcase2023_task4_baseline/local/
    sed_trainer.py (parsed with ast: the recipe module itself needs Lightning); the mirror may only ADD **kwargs."""
    from desed_task_b200.sed_trainer import SEDTask4
    ref = CONTRACT["SEDTask4"][name]
    ps = list(inspect.signature(getattr(SEDTask4, name)).parameters.values())
    pos = [p for p in ps if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
    got = [[p.name, None if p.default is inspect.Parameter.empty else repr(p.default)] for p in pos]
    assert got == ref["args"], name
    if ref["vararg"]:
        assert any(p.kind == p.VAR_POSITIONAL for p in ps)
    if ref["kwarg"]:
        assert any(p.kind == p.VAR_KEYWORD for p in ps)


# every `desed_task` import statement of the four recipe files that drive the hot path (recipes/dcase2023_task4_baseline/
# train_sed.py:11-16, local/sed_trainer.py:14-18, recipes/dcase2024_task4_baseline/train_pretrained.py:21-26,
# local/sed_trainer_pretrained.py:20-25); test_recipe_import_list_is_current re-derives the list from the live reference
RECIPE_IMPORTS = [
    "from desed_task.dataio import ConcatDatasetBatchSampler",
    "from desed_task.dataio.datasets import StronglyAnnotatedSet, UnlabeledSet, WeakSet",
    "from desed_task.nnet.CRNN import CRNN",
    "from desed_task.utils.encoder import ManyHotEncoder",
    "from desed_task.utils.encoder import CatManyHotEncoder, ManyHotEncoder",
    "from desed_task.utils.schedulers import ExponentialWarmup",
    "from desed_task.data_augm import mixup",
    "from desed_task.evaluation.evaluation_measures import compute_per_intersection_macro_f1, compute_psds_from_operating_points, compute_psds_from_scores",
    "from desed_task.utils.postprocess import ClassWiseMedianFilter",
    "from desed_task.utils.scaler import TorchScaler",
]
RECIPE_FILES = ["recipes/dcase2023_task4_baseline/train_sed.py", "recipes/dcase2023_task4_baseline/local/sed_trainer.py",
                "recipes/dcase2024_task4_baseline/train_pretrained.py",
                "recipes/dcase2024_task4_baseline/local/sed_trainer_pretrained.py"]
# third-party packages of the reference's own requirements that this image does not have
ABSENT_OK = {"h5py", "dcase_util", "psds_eval", "sed_eval", "sed_scores_eval"}
HOT = {"desed_task.nnet.CRNN", "desed_task.data_augm", "desed_task.utils.scaler", "desed_task.utils.postprocess",
       "desed_task.utils.schedulers"}


def _reference_root():
    for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "desed_task", "dataio")):
            return cand
    return None


_IMPORT_CACHE = {}


def _import_results(ref, stub):
    """One child process imports every statement (each in a fresh sub-interpreter-like state would cost ~4 s of `import
    torch` apiece): per statement -> ("FILE", path) | ("ABSENT", module) | ("ERROR", message)."""
    key = (ref, stub)
    if key in _IMPORT_CACHE:
        return _IMPORT_CACHE[key]
    code = ["import sys, importlib, json", "sys.path[:0] = [%r, %r]" % (ROOT, ref)]
    if stub:
        code += ["from unittest.mock import MagicMock",
                 "for n in %r + ['dcase_util.data']: sys.modules[n] = MagicMock()" % sorted(ABSENT_OK)]
    code += ["out = {}", "for stmt in %r:" % RECIPE_IMPORTS,
             "    try:", "        exec(stmt, {})",
             "        out[stmt] = ['FILE', importlib.import_module(stmt.split()[1]).__file__]",
             "    except ModuleNotFoundError as e:", "        out[stmt] = ['ABSENT', e.name]",
             "    except Exception as e:", "        out[stmt] = ['ERROR', repr(e)]",
             "    for k in [k for k in sys.modules if k.startswith('desed_task.') and k.split('.')[1] in "
             "('dataio', 'evaluation')]:", "        del sys.modules[k]",      # a failed import must not poison the next
             "print('RESULT ' + json.dumps(out))"]
    r = subprocess.run([sys.executable, "-c", "\n".join(code)], capture_output=True, text=True, cwd="/")
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    _IMPORT_CACHE[key] = json.loads(line[7:])
    return _IMPORT_CACHE[key]


@pytest.mark.parametrize("stmt", RECIPE_IMPORTS)
def test_recipe_import_blocks_resolve(stmt):
    """With this repository in front of the reference on sys.path, EVERY desed_task import of the recipes resolves: hot-path
    names to desed_task_b200, everything else (dataio, encoder, evaluation) to the reference's own files."""
    ref = _reference_root()
    if ref is None:
        if stmt.split()[1] not in HOT:
            pytest.skip("no reference checkout here (the non-hot modules live in the reference)")
        ref = "/nonexistent"
    mod = stmt.split()[1]
    tag, val = _import_results(ref, False)[stmt]
    if tag == "ABSENT":
        assert val in ABSENT_OK, (stmt, val)              # only a third-party module of the reference may be missing
        tag, val = _import_results(ref, True)[stmt]       # ... and with that dependency stubbed the import completes
    assert tag == "FILE", (stmt, tag, val)
    if mod in HOT:
        assert val.startswith(os.path.join(ROOT, "desed_task") + os.sep), val
    else:
        assert val.startswith(ref + os.sep), val


def test_hot_path_classes_are_the_b200_ones():
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from desed_task.nnet.CRNN import CRNN\n"
            "from desed_task.utils import TorchScaler, ExponentialWarmup\n"
            "import desed_task_b200.nnet.CRNN as m; assert CRNN is m.CRNN\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/")
    assert out.returncode == 0, out.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir("/root/reference/recipes"), reason="live reference only in the build container")
def test_recipe_import_list_is_current():
    import ast
    found = set()
    for f in RECIPE_FILES:
        for node in ast.parse(open(os.path.join("/root/reference", f)).read()).body:
            if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("desed_task"):
                found.add("from %s import %s" % (node.module, ", ".join(a.name for a in node.names)))
            elif isinstance(node, ast.Import):
                assert not any(a.name.startswith("desed_task") for a in node.names)
    assert found == set(RECIPE_IMPORTS), found ^ set(RECIPE_IMPORTS)


@pytest.mark.skipif(not os.path.isdir("/root/reference/desed_task"), reason="live reference only in the build container")
def test_fixture_is_current_against_the_live_reference():
    out = subprocess.run([sys.executable, "-c",
                          "import sys, json; sys.path.insert(0, %r); import make_api_contract as m; "
                          "print(json.dumps(m.contract(), sort_keys=True))" % os.path.join(ROOT, "oracle")],
                         capture_output=True, text=True, cwd="/")
    assert out.returncode == 0, out.stderr[-2000:]
    assert json.loads(out.stdout.strip().splitlines()[-1]) == json.loads(json.dumps(CONTRACT, sort_keys=True))


def test_constructor_quirks_of_the_reference_are_kept():
    """Probe-verified behaviours of desed_task.nnet.CRNN that recipes (silently) rely on - SURVEY.md section 7 'API quirks'."""
    from desed_task_b200.nnet.CRNN import CRNN
    cfg = dict(CONTRACT["nets"]["2024"]["config"])
    assert cfg.get("rnn_layers") == 1                       # the YAML key is `rnn_layers`, the ctor argument `n_layers_RNN`
    net = CRNN(**cfg)
    assert net.rnn.rnn.num_layers == 2                      # ... so the 2024 config still builds a 2-layer GRU (CRNN.py:22,37)
    assert net.train() is None and net.eval() is None       # CRNN.train returns None (CRNN.py:308-323)
    assert net.training is False
    with pytest.raises(AttributeError):
        CRNN(nclass=(10, 17))                               # multi-head + attention crashes in the ctor (CRNN.py:113)
    CRNN(rnn_type="BLSTM")                                  # a non-BGRU rnn_type does not raise at construction (CRNN.py:101)
    assert CRNN(**dict(cfg, median_filter=[3] * 27, some_future_key=1)) is not None     # unknown keys are dropped (CNN.py:45)


def test_no_cpu_fallback():
    """The product path fails loudly without a CUDA device: no PyTorch / oracle arithmetic behind the module API."""
    from desed_task_b200._lib import SedkError
    from desed_task_b200.data_augm import mixup
    from desed_task_b200.nnet.CRNN import CRNN
    from desed_task_b200.utils.scaler import TorchScaler
    net = CRNN(**CONTRACT["nets"]["2023"]["config"])
    with pytest.raises(SedkError):
        net(torch.zeros(1, 128, 626))
    with pytest.raises(SedkError):
        mixup(torch.zeros(4, 128, 626), torch.zeros(4, 10, 156))
    with pytest.raises(SedkError):
        TorchScaler("instance", "minmax", [1, 2])(torch.zeros(2, 128, 626))
    import desed_task_b200
    src = "".join(open(os.path.join(os.path.dirname(desed_task_b200.__file__), f)).read()
                  for f in ("engine.py", "frontend.py", "data_augm.py", "optim.py", "sed_trainer.py", "nnet/CRNN.py"))
    assert "import oracle" not in src and "from oracle" not in src


def test_modules_deepcopy_and_pickle_after_use():
    """The reference modules are freely picklable / deep-copyable; the kernel-side caches (ctypes structures full of device
    pointers) must not leak into a copy (deepcopy(SEDTask4), torch.save(module), Lightning ddp_spawn)."""
    import copy
    import ctypes
    import pickle
    from desed_task_b200._lib import CrnnPlan, MelTables
    from desed_task_b200.frontend import MelSpectrogram
    from desed_task_b200.nnet.CRNN import CRNN
    mel = MelSpectrogram(16000, 2048, 2048, 256, 0, 8000, n_mels=128, window_fn=torch.hamming_window,
                         wkwargs={"periodic": False}, power=1)
    mel._tables = {"tab": MelTables(window=ctypes.c_void_p(1234)), "key": ("cuda:0",)}       # what a first run() leaves
    net = CRNN(**CONTRACT["nets"]["2023"]["config"])
    net._ws = {("k",): CrnnPlan()}
    for m in (mel, net):
        for clone in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
            assert [k for k, _ in clone.state_dict().items()] == [k for k, _ in m.state_dict().items()]
            assert not getattr(clone, "_tables", None) and not getattr(clone, "_ws", None)
    # seeds are reproducible across runs: the per-instance salt is a construction counter, not an address
    a, b = CRNN(**CONTRACT["nets"]["2023"]["config"]), CRNN(**CONTRACT["nets"]["2023"]["config"])
    assert b._instance == a._instance + 1 and copy.deepcopy(a)._instance > b._instance


def test_configuration_coverage_is_explicit():
    """What the kernels cover is accepted, everything else is named by _unsupported() (and raises at the first forward): no
    configuration silently computes something else than the reference."""
    from desed_task_b200.nnet.CRNN import CRNN
    c23, c24 = CONTRACT["nets"]["2023"]["config"], CONTRACT["nets"]["2024"]["config"]
    for cfg in (c23, c24, dict(c23, dropstep_recurrent=0.3), dict(c24, dropstep_recurrent=0.3),
                dict(c24, aggregation_type="interpolate"), dict(c23, activation="cg"), dict(c23, activation="Relu"),
                dict(c23, activation="leakyrelu"), dict(c23, freeze_bn=True), dict(c23, train_cnn=False)):
        assert CRNN(**cfg)._unsupported() is None, cfg
    for cfg, word in ((dict(c24, aggregation_type="frame"), "aggregation_type"), (dict(c23, normalization="layer"), "normalization"),
                      (dict(c23, rnn_type="BLSTM"), "rnn_type"), (dict(c23, attention=False), "attention"),
                      (dict(c23, n_RNN_cell=100), "n_RNN_cell"), (dict(c23, dropout_recurrent=0.1), "dropout_recurrent"),
                      (dict(c23, activation="tanh"), "activation")):
        assert word in CRNN(**cfg)._unsupported(), cfg
