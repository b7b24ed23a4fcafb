"""Mint tests/golden/api_contract.json: the drop-in contract of the hot path as the LIVE reference defines it.

Test infrastructure only; run in the build container (needs /root/reference):
    python oracle/make_api_contract.py
Records, from the reference's own modules: the constructor / function signatures the recipes call
(`CRNN`, `CNN`, `BidirectionalGRU`, `mixup`, `frame_shift`, `add_noise`, `TorchScaler`, `ExponentialWarmup`,
`ClassWiseMedianFilter`) and, for the shipped 2023 and 2024 `net` configs, the `state_dict` keys / shapes and the
`parameters()` order (relied on by `zip(ema.parameters(), model.parameters())`, sed_trainer.py:198, and by published
checkpoints).  tests/test_dropin_cpu.py holds the host mirror to this file (and this file to the live reference when present).
"""
import importlib.util
import inspect
import json
import os
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "api_contract.json")


def _load(path, name):
    s = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(s)
    s.loader.exec_module(m)
    return m


def sig(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        out.append([p.name, str(p.kind), None if p.default is inspect.Parameter.empty else repr(p.default)])
    return out


def ast_methods(path, cls, names):
    """Signatures of methods of a class whose module cannot be imported here (the recipe trainers need Lightning)."""
    import ast
    out = {}
    for node in ast.parse(open(path).read()).body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name in names:
                    a = f.args
                    pos = [x.arg for x in a.args]
                    dfl = [None] * (len(pos) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
                    out[f.name] = {"args": [[n, d] for n, d in zip(pos, dfl)], "vararg": a.vararg.arg if a.vararg else None,
                                   "kwarg": a.kwarg.arg if a.kwarg else None}
    return out


def net_config(path):
    import yaml
    return yaml.safe_load(open(path))["net"]


def contract():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from baseline import refload                       # private package name: never the repo's own `desed_task` shim
    R = refload.load(REF)
    CRNN = R.CRNN
    CNN, GLU, ContextGating = R.CNN.CNN, R.CNN.GLU, R.CNN.ContextGating
    BidirectionalGRU = R.RNN.BidirectionalGRU
    add_noise, frame_shift, mixup = R.data_augm.add_noise, R.data_augm.frame_shift, R.data_augm.mixup
    scaler, sched, post = R.scaler, R.schedulers, R.postprocess
    c = {"signatures": {
        "CRNN.__init__": sig(CRNN.__init__), "CRNN.forward": sig(CRNN.forward), "CNN.__init__": sig(CNN.__init__),
        "GLU.__init__": sig(GLU.__init__), "ContextGating.__init__": sig(ContextGating.__init__),
        "BidirectionalGRU.__init__": sig(BidirectionalGRU.__init__), "mixup": sig(mixup), "frame_shift": sig(frame_shift),
        "add_noise": sig(add_noise), "TorchScaler.__init__": sig(scaler.TorchScaler.__init__),
        "ExponentialWarmup.__init__": sig(sched.ExponentialWarmup.__init__),
        "ClassWiseMedianFilter.__init__": sig(post.ClassWiseMedianFilter.__init__)}, "nets": {}}
    for tag, path in (("2023", f"{REF}/recipes/dcase2023_task4_baseline/confs/default.yaml"),
                      ("2024", f"{REF}/recipes/dcase2024_task4_baseline/confs/pretrained.yaml")):
        cfg = net_config(path)
        net = CRNN(**cfg)
        c["nets"][tag] = {"config": cfg,
                          "state_dict": [[k, list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()],
                          "parameters": [n for n, _ in net.named_parameters()],
                          "n_params": sum(p.numel() for p in net.parameters())}
    # the Lightning module's hot-path surface (recipes/dcase2023_task4_baseline/local/sed_trainer.py:41-55,163,187,253,266,269,358)
    c["SEDTask4"] = ast_methods(f"{REF}/recipes/dcase2023_task4_baseline/local/sed_trainer.py", "SEDTask4",
                                ("__init__", "training_step", "update_ema", "detect", "take_log", "lr_scheduler_step",
                                 "on_before_zero_grad", "configure_optimizers"))
    return c


if __name__ == "__main__":
    c = contract()
    json.dump(c, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, {k: v["n_params"] for k, v in c["nets"].items()})
