"""Static checks that need neither a GPU nor the built library: the product package never touches the oracle,
every C-ABI name the Python wrappers call is declared in include/xmodal_b200.h, and the CPU stand-ins of the
device ops (tests/fake_ops.py) keep the signatures of the real wrappers they replace."""
import ast
import inspect
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "multimodal_eeg_fmri_b200"


def test_product_package_never_imports_the_oracle():
    for path in PKG.glob("*.py"):
        tree = ast.parse(path.read_text())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n == "oracle" or n.startswith("oracle.") for n in names), f"{path.name} imports the oracle"
    for path in (PKG / "csrc").glob("*"):
        assert "oracle" not in path.read_text().lower() or path.suffix not in (".cu", ".cuh"), path.name


def test_every_called_entry_point_is_declared_in_the_header():
    header = (ROOT / "include" / "xmodal_b200.h").read_text()
    declared = set(re.findall(r"\b(xm_[a-z0-9_]+)\s*\(", header))
    called = set()
    for path in PKG.glob("*.py"):
        called |= set(re.findall(r"[\"'](xm_[a-z0-9_]+)[\"']", path.read_text()))
        called |= set(re.findall(r"lib\(\)\.(xm_[a-z0-9_]+)", path.read_text()))
    missing = sorted(called - declared)
    assert not missing, f"called but not declared in include/xmodal_b200.h: {missing}"
    # and every declared compute entry point has a definition in csrc/
    src = "\n".join(p.read_text() for p in (PKG / "csrc").glob("*.cu"))
    undefined = sorted(n for n in declared if not re.search(r"\b" + n + r"\s*\(", src))
    assert not undefined, f"declared but not defined: {undefined}"


def test_fake_ops_keep_the_signatures_of_the_real_wrappers():
    import importlib.util
    import sys
    sys.path.insert(0, str(ROOT / "tests"))
    import fake_ops  # noqa: E402
    spec = importlib.util.spec_from_file_location("_xm_ops_src", PKG / "ops.py")
    real_src = ast.parse((PKG / "ops.py").read_text())
    real = {n.name: n for n in real_src.body if isinstance(n, ast.FunctionDef)}
    assert spec is not None
    checked = 0
    for name, fn in inspect.getmembers(fake_ops, inspect.isfunction):
        if name.startswith("_") or name not in real:
            continue
        fake_params = list(inspect.signature(fn).parameters)
        real_params = [a.arg for a in real[name].args.args]
        # the fake may omit trailing keyword-only tuning knobs of the real wrapper, never reorder or rename
        assert real_params[:len(fake_params)] == fake_params or fake_params[:len(real_params)] == real_params, \
            f"{name}: fake {fake_params} vs real {real_params}"
        checked += 1
    assert checked >= 25
