"""CPU suite: the C-ABI library loads and exports every symbol include/fpd_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "fpd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fpd_[a-z0-9_]+)\s*\(", src)) - {"fpd_stream_t"})


def test_library_exports_every_declared_symbol():
    import fpd_b200  # noqa: F401
    from fpd_b200 import _native as N
    assert os.path.exists(N.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    h = ctypes.CDLL(N.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(h, name), "missing export: " + name
    # the ctypes table covers the whole header and nothing else
    assert sorted(N.EXPORTED_SYMBOLS) == declared
    assert N.lib().fpd_version() >= 100
    assert N.lib().fpd_conv2d_tc_supported(64, 64, 3) == 1
    assert N.lib().fpd_conv2d_tc_supported(3, 32, 7) == 0


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    import fpd_b200  # noqa: F401
    from fpd_b200 import _native as N
    monkeypatch.setattr(N, "_lib", None)
    monkeypatch.setattr(N, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(N.NativeLibraryMissing):
        N.lib()


def test_every_host_module_imports_on_cpu():
    """Import (= byte-compile) every Python module of the package and the top-level scripts: no GPU needed."""
    import importlib
    import fpd_b200  # noqa: F401
    for name in ("_native", "ops", "engine", "engine_hrnet", "autograd_bridge", "train_step", "parallel",
                 "lib.models.hourglass", "lib.models.pose_hrnet", "lib.core.loss", "lib.core.function",
                 "lib.core.inference", "lib.core.evaluate", "lib.utils.transforms", "lib.nms.nms"):
        importlib.import_module("fpd_b200." + name)
    import ast
    for script in ("bench.py", "__graft_entry__.py", "tools/profile_step.py", "tools/profile_kernel.py",
                   "tools/profile_convs.py", "tools/bench_conv_variants.py", "tools/diag_net.py", "tools/diag_grad.py",
                   "tools/summarize_profiles.py"):
        ast.parse(open(os.path.join(ROOT, script)).read(), script)
