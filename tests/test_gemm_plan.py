"""CPU test of the GEMM group launch's host-side planning arithmetic (kaldi-lstm_b200/csrc/lstmp_gemm_plan.h: work-item
tiling, split-K plan of a group) through tests/cpp/plan_test.cc -- g++ only, no CUDA, no GPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "plan_test")


def test_gemm_group_plan_on_cpu():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s", "_build/plan_test"])
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=120)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and r.stdout.strip().endswith("PASS"), r.stdout + r.stderr
    # the plans of the shapes the engine launches are part of the output (one line per group)
    assert "cfg3 layer 2 (4 products)" in r.stdout and "cfg5 per GPU" in r.stdout


def test_engine_uses_the_tested_header():
    """lstmp_gemm_hl.cu plans with the header the CPU test covers (no second copy of the arithmetic in the .cu)."""
    src = open(os.path.join(ROOT, "kaldi-lstm_b200", "csrc", "lstmp_gemm_hl.cu")).read()
    assert '#include "lstmp_gemm_plan.h"' in src
    assert "hlplan::plan_group(" in src and "hlplan::tiling(" in src
    assert "group_makespan" not in src
