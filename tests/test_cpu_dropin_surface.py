"""Drop-in surface: every name the reference's example scripts import from `osrl.*` exists in the
alias package (compat/osrl), and every keyword they pass to those classes / functions is accepted by the mirrors.

The scripts are parsed, not executed (they need a GPU for our models): for each `examples/train/*.py` and
`examples/eval/*.py` of the reference the test collects `from osrl... import X` and the calls `X(...)` /
`trainer.method(...)`, and checks them against `inspect.signature` of the mirror.  Skipped where /root/reference is
absent (the GPU box); the executable counterpart is tests/test_gpu_dropin_loop.py."""
import ast
import glob
import inspect
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("OSRL_REFERENCE_ROOT", "/root/reference")
SCRIPTS = sorted(glob.glob(os.path.join(REF, "examples", "train", "train_*.py")) +
                 glob.glob(os.path.join(REF, "examples", "eval", "eval_*.py")))
pytestmark = pytest.mark.skipif(not SCRIPTS, reason="reference tree not available")


@pytest.fixture(scope="module")
def alias():
    for m in [m for m in sys.modules if m == "osrl" or m.startswith("osrl.")]:
        del sys.modules[m]
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    try:
        import osrl.algorithms  # noqa: F401
        import osrl.common  # noqa: F401
        import osrl.common.dataset  # noqa: F401
        import osrl.common.exp_util  # noqa: F401
        import osrl
        assert "compat" in osrl.__file__
        yield osrl
    finally:
        sys.path.remove(os.path.join(ROOT, "compat"))
        for m in [m for m in sys.modules if m == "osrl" or m.startswith("osrl.")]:
            del sys.modules[m]


def _accepts(fn, kwargs, n_pos):
    sig = inspect.signature(fn)
    params = sig.parameters
    if any(p.kind == p.VAR_KEYWORD for p in params.values()):
        names_ok = True
    else:
        names_ok = all(k in params for k in kwargs)
    positional = [p for p in params.values() if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD) and p.name != "self"]
    return names_ok and (n_pos <= len(positional) or any(p.kind == p.VAR_POSITIONAL for p in params.values()))


@pytest.mark.parametrize("script", SCRIPTS, ids=[os.path.basename(s) for s in SCRIPTS])
def test_script_surface(alias, script):
    import importlib
    tree = ast.parse(open(script).read())
    imported = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("osrl"):
            mod = importlib.import_module(node.module)
            for a in node.names:
                assert hasattr(mod, a.name), f"{os.path.basename(script)}: {node.module}.{a.name} missing"
                imported[a.asname or a.name] = getattr(mod, a.name)
    assert imported, "script imports nothing from osrl?"
    trainer_cls = [v for k, v in imported.items() if k.endswith("Trainer")]
    checked = 0
    for node in ast.walk(tree):
        if not isinstance(node, ast.Call):
            continue
        kwargs = [k.arg for k in node.keywords if k.arg]
        if isinstance(node.func, ast.Name) and node.func.id in imported:
            target = imported[node.func.id]
            fn = target.__init__ if inspect.isclass(target) else target
            assert _accepts(fn, kwargs, len(node.args)), \
                f"{os.path.basename(script)}: {node.func.id}({', '.join(kwargs)}) not accepted by {fn}"
            checked += 1
        elif isinstance(node.func, ast.Attribute) and isinstance(node.func.value, ast.Name) and \
                node.func.value.id == "trainer" and trainer_cls:
            assert hasattr(trainer_cls[0], node.func.attr), f"{os.path.basename(script)}: trainer.{node.func.attr} missing"
            assert _accepts(getattr(trainer_cls[0], node.func.attr), kwargs, len(node.args)), \
                f"{os.path.basename(script)}: trainer.{node.func.attr} call not accepted"
            checked += 1
        elif isinstance(node.func, ast.Attribute) and isinstance(node.func.value, ast.Name) and \
                node.func.value.id == "model" and node.func.attr in ("load_state_dict", "state_dict", "to", "eval", "parameters"):
            checked += 1   # nn.Module API: the mirrors are nn.Modules
    assert checked >= 3, f"{os.path.basename(script)}: only {checked} calls recognised"
