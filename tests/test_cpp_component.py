"""The C++ host-side mirror of the reference component (kaldi-lstm_b200/kaldi/b200-lstm-projected-streams.h,
built against the compat Kaldi surface) exercised through its own test driver (tests/cpp/component_test.cc)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "component_test")
ORACLE = os.path.join(ROOT, "oracle", "_build", "liblstmp_oracle.so")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-s"])
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])


def test_cpp_component_builds_and_fails_loudly_without_gpu():
    _build()
    assert os.path.exists(BIN)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([BIN, ORACLE], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_cpp_component_parity_on_gpu():
    _build()
    r = subprocess.run([BIN, ORACLE], capture_output=True, text=True, timeout=120)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_xent_mirror_on_gpu():
    """kaldi-lstm_b200/kaldi/b200-nnet-loss.h (B200Xent::EvalMasked / Report) against a dense host restatement."""
    _build()
    xbin = os.path.join(ROOT, "tests", "cpp", "_build", "xent_test")
    r = subprocess.run([xbin], capture_output=True, text=True, timeout=120)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr



@pytest.mark.gpu
def test_cpp_trainer_loop_on_gpu():
    """tests/cpp/trainer_test.cc: the reference trainer's main loop (bd-nnet-train-lstm-streams.cc:143-229) in C++ over
    B200StreamDispatch + B200LstmProjectedStreams + B200AffineSoftmaxXent -- frame accounting, cross-validation mode,
    decreasing loss."""
    _build()
    tbin = os.path.join(ROOT, "tests", "cpp", "_build", "trainer_test")
    r = subprocess.run([tbin], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
