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


def test_recipe_import_lines_resolve():
    """The import lines of recipes/dcase202{3,4}_task4_baseline/{train_*.py,local/sed_trainer*.py} that belong to the hot path."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from desed_task.nnet.CRNN import CRNN\n"
            "from desed_task.data_augm import mixup\n"
            "from desed_task.utils.scaler import TorchScaler\n"
            "from desed_task.utils.postprocess import ClassWiseMedianFilter\n"
            "from desed_task.utils.schedulers import ExponentialWarmup\n"
            "import desed_task_b200.nnet.CRNN as m; assert CRNN is m.CRNN\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/")
    assert out.returncode == 0, out.stderr[-2000:]


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
