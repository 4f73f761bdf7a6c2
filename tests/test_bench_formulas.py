"""CPU checks of bench.py's host-side arithmetic: the algorithmic bytes / flops behind `roofline` and `chunk_roofline`
are the figures SURVEY.md section 8(d) states, both arms describe the same `config`, and the reference arm's line has
the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def bench():
    import bench as b
    return b


# (I, C, R, S, T) -> (GFLOP per chunk, compulsory MB per chunk), SURVEY.md section 8(d) "ALGORITHMIC bytes / flops"
SURVEY_8D = {
    "cfg2": ((40, 800, 512, 4, 20), 1.044, 41.1),
    "cfg3 layer 1": ((40, 800, 512, 64, 20), 16.71, 134.4),
    "cfg3 layer 2": ((512, 800, 512, 64, 20), 28.31, 163.4),
    "cfg4 per GPU": ((40, 800, 512, 32, 20), 8.36, 84.7),
    "cfg5 per GPU": ((40, 2048, 1024, 64, 20), 83.05, 420.1),
}


@pytest.mark.parametrize("name", sorted(SURVEY_8D))
def test_chunk_bytes_and_flops_are_the_survey_figures(bench, name):
    shape, gflop, mb = SURVEY_8D[name]
    B, F = bench.alg_chunk(*shape)
    assert abs(F / 1e9 - gflop) <= 0.006, (name, F)
    assert abs(B / 1e6 - mb) <= 0.06, (name, B)


def test_flops_per_frame(bench):
    # 13.056 MFLOP per frame at 40/800/512, 22.118 for the stacked layer, 64.881 for cfg5
    for (I, C, R), mflop in (((40, 800, 512), 13.056), ((512, 800, 512), 22.118), ((40, 2048, 1024), 64.881)):
        _, F = bench.alg_chunk(I, C, R, 1, 1)
        assert abs(F / 1e6 - mflop) < 1e-3


def test_time_loop_kernel_bytes(bench):
    """Per-launch bytes of the two time-loop kernels = what each must read and write once."""
    I, C, R, S, T = 40, 800, 512, 64, 20
    ts = S * T
    fb, ff = bench.alg_fwd_kernel(I, C, R, S, T)
    bb, bf = bench.alg_bwd_kernel(I, C, R, S, T)
    weights = 4 * C * R + R * C
    # forward: pre-activations in, g,i,f,o + c,h,m + r (record and out) out, weights once, state in and out
    assert fb == 4 * (ts * 4 * C + weights + S * (C + R) + ts * 4 * C + 3 * ts * C + 2 * ts * R + S * (C + R))
    assert bb == 4 * (ts * 4 * C + (T + 1) * S * C + ts * C + ts * R + weights + ts * 4 * C + ts * R + 7 * C)
    assert ff == bf == ts * (2 * R * 4 * C + 2 * C * R)          # recurrent + projection products: 5.24 GFLOP
    assert round(fb / 1e6, 1) == 59.2 and round(bb / 1e6, 1) == 54.6
    # the per-timestep figures of SURVEY section 8d: 209.7 + 52.4 MFLOP per step at S = 64
    assert abs(ff / T / 1e6 - (209.7 + 52.4)) < 0.1


def test_both_arms_describe_the_same_config(bench):
    class A:
        workload = "cfg3"
    wl = bench.WORKLOADS["cfg3"]
    assert bench.workload_config(A, wl, 1) == bench.workload_config(A, wl, 1)
    c1, c8 = bench.workload_config(A, wl, 1), bench.workload_config(A, wl, 8)
    assert set(c1) == set(c8) and c1["workload"] == c8["workload"] and "model" not in c1
    assert c1["num_stream_per_gpu"] == 64 and c1["bptt_frames"] == 20
    assert "larger than L2" in c1["l2"]


def test_reference_arm_line_has_the_contract_keys():
    """bench.py --impl reference runs here (CPU only): one JSON line with the keys the driver reads."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg2",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["e2e"]["d2h_bytes_per_step"] == 0
