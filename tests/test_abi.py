"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/lstmp_b200.h declares, and fails loudly (no CPU fallback) without a GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lstmp_b200.h")).read()
    return sorted(set(re.findall(r"\b(lstmp_b200_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    import kaldi_lstm_b200 as klb
    L = klb.load_library()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "missing export %s" % n
    from kaldi_lstm_b200 import engine
    assert sorted(engine.ABI_SYMBOLS) == names
    assert L.lstmp_b200_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import kaldi_lstm_b200 as klb
    with pytest.raises(klb.EngineError) as ei:
        klb.Engine(40, 800, 512, 4, 20)
    assert ei.value.code == -2  # LSTMP_B200_ENODEV
    assert "no CPU path" in str(ei.value)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under kaldi-lstm_b200/ may import, link or load it."""
    pkg = os.path.join(ROOT, "kaldi-lstm_b200")
    bad = re.compile(r"(import\s+oracle|from\s+oracle|oracle_py|liblstmp_oracle|lstmp_oracle_f|oracle/)")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, f)).read()
                assert not bad.search(txt), os.path.join(dp, f)


def test_built_for_sm100a_with_tma():
    import shutil
    import subprocess
    import kaldi_lstm_b200 as klb
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", klb.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass  # cp.async.bulk weight staging
