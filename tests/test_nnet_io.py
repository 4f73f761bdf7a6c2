"""Host-only tests of the Kaldi nnet1 text-model reader/writer and the google <-> standard conversion
(kaldi-lstm_b200/nnet_io.py; reference README.md:19-29, Q3)."""
import numpy as np
import pytest

import kaldi_lstm_b200 as klb
from kaldi_lstm_b200 import nnet_io as nio


def _google_model(I=6, C=4, R=3, P=5, S=4, seed=0):
    rng = np.random.RandomState(seed)
    n = 4 * C * I + 4 * C * R + 4 * C + 3 * C + R * C
    flat = rng.randn(n).astype(np.float32) * 0.1
    lstm = nio.lstm_component_from_flat(flat, R, I, C, num_stream=S)
    aff = nio.NnetComponent("<AffineTransform>", P, R, [("<LearnRateCoef>", 1.0), ("<BiasLearnRateCoef>", 1.0),
                                                        ("<MaxNorm>", 0.0)],
                            [rng.randn(P, R).astype(np.float32), rng.randn(P).astype(np.float32)])
    return [nio.NnetComponent("<Transmit>", I, I), lstm, aff, nio.NnetComponent("<Softmax>", P, P)], flat


def test_text_roundtrip_is_exact():
    comps, flat = _google_model()
    text = nio.format_nnet(comps)
    assert text.startswith("<Nnet>") and text.rstrip().endswith("</Nnet>")
    assert "<LstmProjectedStreams> 3 6 <CellDim> 4 <NumStream> 4" in text   # README.md:39 header layout
    back = nio.parse_nnet(text)
    assert [c.type for c in back] == ["<Transmit>", "<LstmProjectedStreams>", "<AffineTransform>", "<Softmax>"]
    np.testing.assert_array_equal(nio.lstm_flat_params(back[1]), flat)       # repr(float32) round-trips bit-exactly
    for a, b in zip(comps[2].arrays, back[2].arrays):
        np.testing.assert_array_equal(a, b)
    assert back[2].attr("<MaxNorm>") == 0.0 and back[1].attr("<NumStream>") == 4
    assert nio.format_nnet(back) == text


def test_google_to_standard_and_back():
    comps, flat = _google_model()
    std = nio.google_to_standard(comps, shift=5)
    assert [c.type for c in std] == ["<TimeShift>", "<LstmProjected>", "<AffineTransform>", "<Softmax>"]
    assert nio.targets_delay(std) == 5 and std[1].attr("<NumStream>") is None and std[1].attr("<CellDim>") == 4
    text = nio.format_nnet(std)
    assert "<TimeShift> 6 6 <Shift> 5" in text and "<LstmProjected> 3 6 <CellDim> 4 " in text   # README.md:24-25
    assert "NumStream" not in text and "Transmit" not in text
    np.testing.assert_array_equal(nio.lstm_flat_params(nio.parse_nnet(text)[1]), flat)           # numbers untouched
    goog = nio.standard_to_google(nio.parse_nnet(text), num_stream=4)
    assert nio.format_nnet(goog) == nio.format_nnet(comps)


def test_shapes_and_param_order_match_the_component():
    names = [n for n, _ in nio.lstm_shapes(512, 40, 800)]
    assert names == ["w_gifo_x", "w_gifo_r", "bias", "peephole_i_c", "peephole_f_c", "peephole_o_c", "w_r_m"]
    n = sum(int(np.prod(s)) for _, s in nio.lstm_shapes(512, 40, 800))
    assert n == 2181600                                                      # 8.726 MB
    # the oracle's slices (GetParams order, LPS.h:162-189) agree with the file order
    from oracle import oracle_py
    sl = oracle_py.param_slices(40, 800, 512)
    off = 0
    for name, shape in nio.lstm_shapes(512, 40, 800):
        a, b = sl[name][0], sl[name][1]
        assert tuple(sl[name][2]) == tuple(shape)
        assert a == off and b - a == int(np.prod(shape)), (name, a, b, off)
        off = b


def test_errors():
    with pytest.raises(RuntimeError):
        nio.parse_nnet("<Nnet> <Gizmo> 3 3 </Nnet>")
    with pytest.raises(RuntimeError):   # ExpectToken <NumStream>
        nio.parse_nnet("<Nnet> <LstmProjectedStreams> 1 1 <CellDim> 1 [ 0 0 0 0 ] </Nnet>")
    with pytest.raises(RuntimeError):   # wrong element count for the header's shape
        nio.parse_nnet("<Nnet> <AffineTransform> 2 2 [ 1 2 3 ] [ 0 0 ] </Nnet>")
    with pytest.raises(RuntimeError):
        nio.lstm_component_from_flat(np.zeros(7, np.float32), 1, 1, 1)


def test_time_shift_rows():
    # standard/nnet/nnet-time-shift.h:42-51: src = clamp(dst + shift, 0, num_frames - 1)
    np.testing.assert_array_equal(nio.time_shift_rows(6, 2), [2, 3, 4, 5, 5, 5])
    np.testing.assert_array_equal(nio.time_shift_rows(4, -1), [0, 0, 1, 2])
    np.testing.assert_array_equal(nio.time_shift_rows(3, 0), [0, 1, 2])
    assert klb.nnet_io is nio


class _FakeEngine:
    """Stands in for the CUDA engine so that the component's model-file logic can be exercised on a CPU-only box."""

    def __init__(self, I, C, R, S, T, device=0):
        self.shape = (I, C, R, S, T)
        self.num_params = 4 * C * I + 4 * C * R + 4 * C + 3 * C + R * C
        self.flat = np.zeros(self.num_params, np.float32)

    def set_flat(self, which, a):
        assert which == 0 and len(a) == self.num_params
        self.flat = np.asarray(a, np.float32).copy()

    def get_flat(self, which):
        return self.flat.copy()


def test_component_model_file_roundtrip(monkeypatch):
    from kaldi_lstm_b200 import component
    monkeypatch.setattr(component, "Engine", _FakeEngine)
    comps, flat = _google_model()
    layer = klb.LstmProjectedStreams.FromNnetComponent(comps[1], max_frames=7)
    assert (layer.input_dim_, layer.ncell_, layer.nrecur_, layer.nstream_) == (6, 4, 3, 4)
    assert layer.engine.shape == (6, 4, 3, 4, 7)
    np.testing.assert_array_equal(layer.GetParams(), flat)
    assert nio.format_nnet([layer.ToNnetComponent()]) == nio.format_nnet([comps[1]])
    # the standard version's component runs as one stream (S = 1) unless told otherwise
    std = nio.google_to_standard(comps, 5)
    one = klb.LstmProjectedStreams.FromNnetComponent(std[1])
    assert one.nstream_ == 1 and one.engine.shape[3] == 1
    np.testing.assert_array_equal(one.GetParams(), flat)
    four = klb.LstmProjectedStreams.FromNnetComponent(std[1], num_stream=4)
    assert four.ToNnetComponent().attr("<NumStream>") == 4
    with pytest.raises(RuntimeError):
        klb.LstmProjectedStreams.FromNnetComponent(comps[0])



# ---- binary files ------------------------------------------------------------------------------------------------
def test_binary_roundtrip_and_text_equivalence(tmp_path):
    comps, flat = _google_model()
    data = nio.format_nnet_binary(comps)
    assert data[:2] == b"\0B" and data[2:9] == b"<Nnet> " and data.endswith(b"</Nnet> ")
    # header of the LSTM component: marker, then output_dim and input_dim as size byte 4 + little-endian int32
    k = data.index(b"<LstmProjectedStreams> ") + len(b"<LstmProjectedStreams> ")
    assert data[k:k + 10] == b"\x04\x03\x00\x00\x00\x04\x06\x00\x00\x00"
    assert data[k + 10:k + 20] == b"<CellDim> " and data[k + 20:k + 25] == b"\x04\x04\x00\x00\x00"
    back = nio.parse_nnet_binary(data)
    assert [c.type for c in back] == ["<Transmit>", "<LstmProjectedStreams>", "<AffineTransform>", "<Softmax>"]
    np.testing.assert_array_equal(nio.lstm_flat_params(back[1]), flat)
    assert back[2].attrs == comps[2].attrs and back[1].attrs == comps[1].attrs
    assert nio.format_nnet_binary(back) == data
    assert nio.format_nnet(back) == nio.format_nnet(comps)                  # same model in text
    # files: the mode is detected from the first two bytes, as Kaldi's Input does
    pb, pt = tmp_path / "final.nnet", tmp_path / "final.nnet.txt"
    nio.write_nnet(str(pb), comps, binary=True)
    nio.write_nnet(str(pt), comps, binary=False)
    for p in (pb, pt):
        rd = nio.read_nnet(str(p))
        np.testing.assert_array_equal(nio.lstm_flat_params(rd[1]), flat)
        np.testing.assert_array_equal(rd[2].arrays[0], comps[2].arrays[0])
    # google -> standard on a binary model (README.md:19-29 does it on text after nnet-copy --binary=false)
    std = nio.google_to_standard(nio.read_nnet(str(pb)), shift=5)
    again = nio.parse_nnet_binary(nio.format_nnet_binary(std))
    assert [c.type for c in again] == ["<TimeShift>", "<LstmProjected>", "<AffineTransform>", "<Softmax>"]
    assert again[0].attr("<Shift>") == 5 and again[1].attr("<NumStream>") is None


def test_binary_reader_tolerates_end_of_component_and_rejects_garbage():
    comps, flat = _google_model()
    data = nio.format_nnet_binary(comps, header=False)
    # later Kaldi versions close every component with <!EndOfComponent>
    marked = data.replace(b"<Softmax> ", b"<!EndOfComponent> <Softmax> ")
    assert [c.type for c in nio.parse_nnet_binary(marked)] == [c.type for c in comps]
    with pytest.raises(RuntimeError, match="Expected token <Nnet>"):
        nio.parse_nnet_binary(b"\0B<Nnot> ")
    with pytest.raises(RuntimeError, match="past the end|truncated|ReadToken"):
        nio.parse_nnet_binary(data[:len(data) // 2])
    bad = data.replace(b"FM ", b"DM ", 1)                                    # a double matrix: not what nnet1 writes
    with pytest.raises(RuntimeError, match="Expected token FM"):
        nio.parse_nnet_binary(bad)
    k = data.index(b"<CellDim> ") + len(b"<CellDim> ")
    wrong = data[:k] + b"\x04\x05\x00\x00\x00" + data[k + 5:]              # CellDim 5: shapes no longer match
    with pytest.raises(RuntimeError, match="component header implies"):
        nio.parse_nnet_binary(wrong)
    with pytest.raises(RuntimeError, match="size byte"):
        nio.parse_nnet_binary(data[:k] + b"\x08" + data[k + 1:])


# ---- the feature transform in front of the network (google/feature_transform.nnet.txt) ------------------------------
def _transform_text(shift, scale, coef=False):
    def vec(v):
        return " [ " + " ".join("%.6f" % x for x in v) + " ]\n"
    lr = "<LearnRateCoef> 0 " if coef else ""
    d = len(shift)
    return "<Nnet> \n<AddShift> %d %d %s\n%s<Rescale> %d %d %s\n%s</Nnet> \n" % (d, d, lr, vec(shift), d, d, lr, vec(scale))


@pytest.mark.parametrize("coef", [False, True])
def test_feature_transform_is_parsed_into_the_dispatchers_two_vectors(coef):
    rng = np.random.RandomState(3)
    shift = (-17 + rng.randn(40)).astype(np.float32)
    scale = (0.25 + 0.01 * rng.rand(40)).astype(np.float32)
    comps = nio.parse_nnet(_transform_text(shift, scale, coef))
    assert [c.type for c in comps] == ["<AddShift>", "<Rescale>"]
    sh, sc = nio.feature_transform(comps)
    np.testing.assert_allclose(sh, shift, atol=1e-6)
    np.testing.assert_allclose(sc, scale, atol=1e-6)
    # binary round trip of the same two components
    back = nio.parse_nnet_binary(nio.format_nnet_binary(comps))
    sh2, sc2 = nio.feature_transform(back)
    np.testing.assert_array_equal(sh2, sh)
    np.testing.assert_array_equal(sc2, sc)
    # what DeviceStreamDispatcher(shift=sh, scale=sc) fuses into its gather kernel: (x + shift) * scale, i.e. raw
    # log-mel features ~ N(17, 3.6^2) come out normalised
    x = (rng.randn(1000, 40) * 3.6 + 17).astype(np.float32)
    y = (x + sh) * sc
    assert y.dtype == np.float32 and abs(y.mean()) < 0.5 and 0.5 < y.std() < 1.5


def test_feature_transform_rejects_what_it_cannot_fuse():
    d = 4
    a = nio.NnetComponent("<AddShift>", d, d, [], [np.zeros(d, np.float32)])
    r = nio.NnetComponent("<Rescale>", d, d, [], [np.ones(d, np.float32)])
    assert nio.feature_transform([]) == (None, None)
    assert nio.feature_transform([r])[0] is None and nio.feature_transform([a])[1] is None
    for bad in ([r, a], [a, a], [a, r, r], [nio.NnetComponent("<Softmax>", d, d)],
                [a, nio.NnetComponent("<Rescale>", 5, 5, [], [np.ones(5, np.float32)])]):
        with pytest.raises(RuntimeError, match="feature transform"):
            nio.feature_transform(bad)


def test_the_references_own_transform_file():
    """The file the reference ships (read where it lies; absent on the GPU box)."""
    import os
    path = "/root/reference/google/feature_transform.nnet.txt"
    if not os.path.exists(path):
        pytest.skip("reference tree absent")
    comps = nio.read_nnet(path)
    sh, sc = nio.feature_transform(comps)
    assert sh.shape == sc.shape == (40,)
    assert -20 < sh.min() and sh.max() < -15 and 0.2 < sc.min() and sc.max() < 0.35   # shift -17..-16, scale 0.25..0.30
