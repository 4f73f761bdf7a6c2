"""Kaldi nnet1 model files (text AND binary) for the components on the recipe's path, and the google <-> standard
conversion.

Host-only (no GPU): SURVEY.md section 8(f) rank 4.  The reference describes the conversion as manual text editing
(README.md:19-29 and Q3): `nnet-copy --binary=false`, change `<Transmit>` to `<TimeShift> ... <Shift> k` (k = the
`--targets-delay` used in training), change `<LstmProjectedStreams>` to `<LstmProjected>` and drop the `<NumStream>` tag.
Both LSTM components write the same seven parameter blocks in the same order (`w_gifo_x, w_gifo_r, bias, peephole_i_c,
peephole_f_c, peephole_o_c, w_r_m`: google/nnet/bd-nnet-lstm-projected-streams.h:133-150,
standard/nnet/nnet-lstm-projected.h:139-152), so the conversion never touches the numbers.

Text layout (upstream nnet-nnet.cc / nnet-component.cc, kaldi-matrix.cc:1172-1211): `<Nnet>`, then per component
`<Type> output_dim input_dim` followed by the component's own data, then `</Nnet>`.  Matrices are ` [` rows `]`,
vectors ` [ a b c ]`; reading is whitespace-tokenised and the shapes come from the component header.

Binary layout (what `nnet-initialize` / the trainer write by default; upstream io-funcs + kaldi-matrix.cc:1172-1211, token
order of the LSTM from LPS.h:101-150): the file starts with `\0B`; tokens are `<Token>` + one space in both modes; an
int32 / float is one size byte (4) + 4 little-endian bytes; a float matrix is `FM ` + int32 rows + int32 cols + the
rows back to back, a float vector `FV ` + int32 dim + data.  `parse_nnet_binary` / `format_nnet_binary` are pinned against
the reference's own ReadData / WriteData compiled here (tests/test_ref_pin.py).
"""
import struct

import numpy as np

LSTM_TYPES = ("<LstmProjectedStreams>", "<LstmProjected>")
TRANSFORM_TYPES = ("<AddShift>", "<Rescale>")   # the feature transform in front of the network (CMVN)


class NnetComponent:
    def __init__(self, type_, output_dim, input_dim, attrs=None, arrays=None):
        self.type = type_                # marker incl. the angle brackets
        self.output_dim = int(output_dim)
        self.input_dim = int(input_dim)
        self.attrs = list(attrs or [])   # ordered (token, value) pairs that follow the dims
        self.arrays = list(arrays or []) # numpy float32 arrays in file order

    def attr(self, token, default=None):
        for k, v in self.attrs:
            if k == token:
                return v
        return default

    def __repr__(self):
        return "NnetComponent(%s %d %d %s, %d arrays)" % (self.type, self.output_dim, self.input_dim, self.attrs,
                                                        len(self.arrays))


def lstm_shapes(output_dim, input_dim, ncell):
    """Parameter blocks of both LSTM components in file / GetParams order (LPS.h:133-150, :162-189)."""
    C, R, I = int(ncell), int(output_dim), int(input_dim)
    return [("w_gifo_x", (4 * C, I)), ("w_gifo_r", (4 * C, R)), ("bias", (4 * C,)), ("peephole_i_c", (C,)),
            ("peephole_f_c", (C,)), ("peephole_o_c", (C,)), ("w_r_m", (R, C))]


class _Tokens:
    def __init__(self, text):
        self.t = text.split()
        self.i = 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def next(self):
        if self.i >= len(self.t):
            raise RuntimeError("unexpected end of nnet text")
        self.i += 1
        return self.t[self.i - 1]

    def expect(self, tok):
        got = self.next()
        if got != tok:
            raise RuntimeError("Expected token %s, got %s" % (tok, got))  # ExpectToken

    def array(self, shape):
        self.expect("[")
        n = int(np.prod(shape))
        vals = self.t[self.i:self.i + n]
        if len(vals) != n or "]" in vals:
            raise RuntimeError("matrix/vector in nnet text has the wrong number of elements for shape %s" % (shape,))
        self.i += n
        self.expect("]")
        return np.array(vals, dtype=np.float64).astype(np.float32).reshape(shape)


def parse_nnet(text):
    """Kaldi nnet1 text model -> list of NnetComponent.  Knows the component types of the recipe
    (google/nnet.proto, standard/nnet.proto): Transmit, TimeShift, LstmProjectedStreams, LstmProjected,
    AffineTransform, Softmax."""
    tk = _Tokens(text)
    tk.expect("<Nnet>")
    comps = []
    while True:
        typ = tk.next()
        if typ == "</Nnet>":
            break
        out_dim, in_dim = int(tk.next()), int(tk.next())
        c = NnetComponent(typ, out_dim, in_dim)
        if typ in ("<Transmit>", "<Softmax>"):
            pass
        elif typ == "<TimeShift>":
            tk.expect("<Shift>")                                      # nnet-time-shift.h:33-36
            c.attrs.append(("<Shift>", int(tk.next())))
        elif typ in LSTM_TYPES:
            tk.expect("<CellDim>")                                    # LPS.h:101-104 / nnet-lstm-projected.h:111-113
            ncell = int(tk.next())
            c.attrs.append(("<CellDim>", ncell))
            if typ == "<LstmProjectedStreams>":
                tk.expect("<NumStream>")
                c.attrs.append(("<NumStream>", int(tk.next())))
            for _, shape in lstm_shapes(out_dim, in_dim, ncell):
                c.arrays.append(tk.array(shape))
        elif typ == "<AffineTransform>":
            while tk.peek() in ("<LearnRateCoef>", "<BiasLearnRateCoef>", "<MaxNorm>"):
                k = tk.next()
                c.attrs.append((k, float(tk.next())))
            c.arrays.append(tk.array((out_dim, in_dim)))
            c.arrays.append(tk.array((out_dim,)))
        elif typ in TRANSFORM_TYPES:                                  # google/feature_transform.nnet.txt:2-5
            if tk.peek() == "<LearnRateCoef>":                        # (later Kaldi versions; absent in the reference's file)
                k = tk.next()
                c.attrs.append((k, float(tk.next())))
            c.arrays.append(tk.array((out_dim,)))
        else:
            raise RuntimeError("Unknown component type %s" % typ)
        comps.append(c)
    return comps


# ---- binary ------------------------------------------------------------------------------------------------------
BINARY_HEADER = b"\0B"
_FLOAT_ATTRS = ("<LearnRateCoef>", "<BiasLearnRateCoef>", "<MaxNorm>")


class _BinReader:
    """Kaldi's binary-mode readers (upstream base/io-funcs-inl.h): ReadToken, ReadBasicType, Matrix / Vector Read."""

    def __init__(self, data):
        self.d = bytes(data)
        self.i = 0

    def _skip_ws(self):
        while self.i < len(self.d) and self.d[self.i:self.i + 1].isspace():
            self.i += 1

    def peek_token(self):
        save = self.i
        try:
            return self.token()
        except RuntimeError:
            return None
        finally:
            self.i = save

    def token(self):
        self._skip_ws()                                    # operator>>(string) skips leading white space
        j = self.i
        while j < len(self.d) and not self.d[j:j + 1].isspace():
            j += 1
        if j == self.i:
            raise RuntimeError("ReadToken failed: unexpected end of binary nnet")
        if j >= len(self.d):
            raise RuntimeError("ReadToken: expected space after token")
        tok = self.d[self.i:j].decode("ascii", "replace")
        self.i = j + 1                                     # exactly one space is consumed: raw bytes may follow
        return tok

    def expect(self, tok):
        got = self.token()
        if got != tok:
            raise RuntimeError("Expected token %s, got %s" % (tok, got))

    def basic(self, fmt):
        if self.i >= len(self.d) or self.d[self.i] != 4:
            raise RuntimeError("ReadBasicType: size byte %r where 4 was expected" % self.d[self.i:self.i + 1])
        if self.i + 5 > len(self.d):
            raise RuntimeError("ReadBasicType failed: truncated file")
        v = struct.unpack_from("<" + fmt, self.d, self.i + 1)[0]
        self.i += 5
        return v

    def int32(self):
        return self.basic("i")

    def float32(self):
        return self.basic("f")

    def _data(self, n):
        nb = 4 * n
        if n < 0 or self.i + nb > len(self.d):
            raise RuntimeError("matrix/vector data runs past the end of the file")
        a = np.frombuffer(self.d, dtype="<f4", count=n, offset=self.i).astype(np.float32)
        self.i += nb
        return a

    def matrix(self, shape=None):
        tok = self.token()
        if tok != "FM":
            raise RuntimeError("Expected token FM, got %s (only float matrices are supported)" % tok)
        r, c = self.int32(), self.int32()
        if shape is not None and (r, c) != tuple(shape):
            raise RuntimeError("matrix is %d x %d, the component header implies %s" % (r, c, (shape,)))
        return self._data(r * c).reshape(r, c)

    def vector(self, shape=None):
        tok = self.token()
        if tok != "FV":
            raise RuntimeError("Expected token FV, got %s (only float vectors are supported)" % tok)
        n = self.int32()
        if shape is not None and (n,) != tuple(shape):
            raise RuntimeError("vector has %d elements, the component header implies %s" % (n, (shape,)))
        return self._data(n)

    def array(self, shape):
        return self.matrix(shape) if len(shape) == 2 else self.vector(shape)


def _read_component_data(c, rd):
    """ReadData of one component from a _BinReader (same grammar as the text branch of parse_nnet)."""
    typ, out_dim, in_dim = c.type, c.output_dim, c.input_dim
    if typ in ("<Transmit>", "<Softmax>"):
        pass
    elif typ == "<TimeShift>":
        rd.expect("<Shift>")
        c.attrs.append(("<Shift>", rd.int32()))
    elif typ in LSTM_TYPES:
        rd.expect("<CellDim>")
        ncell = rd.int32()
        c.attrs.append(("<CellDim>", ncell))
        if typ == "<LstmProjectedStreams>":
            rd.expect("<NumStream>")
            c.attrs.append(("<NumStream>", rd.int32()))
        for _, shape in lstm_shapes(out_dim, in_dim, ncell):
            c.arrays.append(rd.array(shape))
    elif typ == "<AffineTransform>":
        while rd.peek_token() in _FLOAT_ATTRS:
            k = rd.token()
            c.attrs.append((k, float(rd.float32())))
        c.arrays.append(rd.matrix((out_dim, in_dim)))
        c.arrays.append(rd.vector((out_dim,)))
    elif typ in TRANSFORM_TYPES:
        if rd.peek_token() == "<LearnRateCoef>":
            k = rd.token()
            c.attrs.append((k, float(rd.float32())))
        c.arrays.append(rd.vector((out_dim,)))
    else:
        raise RuntimeError("Unknown component type %s" % typ)


def parse_nnet_binary(data):
    """Kaldi nnet1 BINARY model (with or without the leading `\\0B`) -> list of NnetComponent."""
    data = bytes(data)
    if data[:2] == BINARY_HEADER:
        data = data[2:]
    rd = _BinReader(data)
    rd.expect("<Nnet>")
    comps = []
    while True:
        typ = rd.token()
        if typ == "</Nnet>":
            break
        if typ == "<!EndOfComponent>":        # written by later Kaldi versions after every component
            continue
        c = NnetComponent(typ, rd.int32(), rd.int32())
        _read_component_data(c, rd)
        comps.append(c)
    return comps


def _tok(t):
    return t.encode("ascii") + b" "


def _i32(v):
    return b"\x04" + struct.pack("<i", int(v))


def _f32(v):
    return b"\x04" + struct.pack("<f", float(v))


def _array_binary(a):
    a = np.ascontiguousarray(a, dtype="<f4")
    if a.ndim == 1:
        return _tok("FV") + _i32(a.shape[0]) + a.tobytes()
    return _tok("FM") + _i32(a.shape[0]) + _i32(a.shape[1]) + a.tobytes()


def component_data_binary(c):
    """What the component's WriteData(os, binary=true) emits (LPS.h:133-150 for the LSTM)."""
    out = []
    for k, v in c.attrs:
        out.append(_tok(k) + (_f32(v) if k in _FLOAT_ATTRS else _i32(v)))
    out.extend(_array_binary(a) for a in c.arrays)
    return b"".join(out)


def format_nnet_binary(comps, header=True):
    out = [BINARY_HEADER if header else b"", _tok("<Nnet>")]
    for c in comps:
        out.append(_tok(c.type) + _i32(c.output_dim) + _i32(c.input_dim) + component_data_binary(c))
    out.append(_tok("</Nnet>"))
    return b"".join(out)


def read_nnet(path):
    """Model file in either mode (Kaldi's Input::Open test: binary files start with `\\0B`)."""
    with open(path, "rb") as f:
        data = f.read()
    return parse_nnet_binary(data) if data[:2] == BINARY_HEADER else parse_nnet(data.decode("ascii"))


def write_nnet(path, comps, binary=True):
    with open(path, "wb") as f:
        f.write(format_nnet_binary(comps) if binary else format_nnet(comps).encode("ascii"))


# ---- text ----------------------------------------------------------------------------------------------------------
def _fmt(v):
    return repr(float(np.float32(v))) if isinstance(v, (float, np.floating)) else str(v)


def _array_text(a):
    a = np.asarray(a, np.float32)
    if a.ndim == 1:                                                    # VectorBase::Write text: " [ a b c ]\n"
        return " [ " + " ".join(_fmt(x) for x in a) + " ]\n"
    rows = ["  " + " ".join(_fmt(x) for x in r) for r in a]            # MatrixBase::Write text, kaldi-matrix.cc:1172-1211
    return " [\n" + "\n".join(rows) + " ]\n"


def format_nnet(comps):
    out = ["<Nnet> \n"]
    for c in comps:
        head = "%s %d %d " % (c.type, c.output_dim, c.input_dim)
        head += "".join("%s %s " % (k, _fmt(v) if isinstance(v, float) else v) for k, v in c.attrs)
        if not c.arrays:
            head += "\n"
        out.append(head + "".join(_array_text(a) for a in c.arrays))
    out.append("</Nnet> \n")
    return "".join(out)


def google_to_standard(comps, shift):
    """README.md:19-29: Transmit -> TimeShift <Shift> shift (= the --targets-delay used in training);
    LstmProjectedStreams -> LstmProjected without <NumStream>.  Parameters are shared by reference, not copied."""
    out = []
    for c in comps:
        if c.type == "<Transmit>":
            out.append(NnetComponent("<TimeShift>", c.output_dim, c.input_dim, [("<Shift>", int(shift))]))
        elif c.type == "<LstmProjectedStreams>":
            out.append(NnetComponent("<LstmProjected>", c.output_dim, c.input_dim,
                                     [(k, v) for k, v in c.attrs if k != "<NumStream>"], c.arrays))
        else:
            out.append(c)
    return out


def standard_to_google(comps, num_stream):
    """The inverse edit (README.md Q3): TimeShift -> Transmit (the trainer applies the target delay itself,
    bd-nnet-train-lstm-streams.cc:198-202), LstmProjected -> LstmProjectedStreams <NumStream> S."""
    out = []
    for c in comps:
        if c.type == "<TimeShift>":
            out.append(NnetComponent("<Transmit>", c.output_dim, c.input_dim))
        elif c.type == "<LstmProjected>":
            attrs = [(k, v) for k, v in c.attrs] + [("<NumStream>", int(num_stream))]
            out.append(NnetComponent("<LstmProjectedStreams>", c.output_dim, c.input_dim, attrs, c.arrays))
        else:
            out.append(c)
    return out


def feature_transform(comps):
    """(shift, scale) of a feature-transform nnet = `<AddShift>` then `<Rescale>` (google/feature_transform.nnet.txt:2-5),
    the two vectors the device-side stream dispatch fuses into its gather kernel: feat = (x + shift) * scale.  Either
    component may be absent (None); anything else -- another component type, the opposite order, a dimension change --
    is not a transform the dispatcher can fuse and raises."""
    shift = scale = None
    for c in comps:
        if c.type not in TRANSFORM_TYPES or c.input_dim != c.output_dim:
            raise RuntimeError("feature transform: cannot fuse component %s %d %d" % (c.type, c.output_dim, c.input_dim))
        if c.type == "<AddShift>":
            if shift is not None or scale is not None:
                raise RuntimeError("feature transform: expected at most one <AddShift>, before <Rescale>")
            shift = np.asarray(c.arrays[0], np.float32)
        else:
            if scale is not None:
                raise RuntimeError("feature transform: expected at most one <Rescale>")
            scale = np.asarray(c.arrays[0], np.float32)
    if shift is not None and scale is not None and shift.shape != scale.shape:
        raise RuntimeError("feature transform: <AddShift> and <Rescale> dimensions differ")
    return shift, scale


def targets_delay(comps):
    """The <Shift> of the model's TimeShift component (None when there is none)."""
    for c in comps:
        if c.type == "<TimeShift>":
            return c.attr("<Shift>")
    return None


def lstm_flat_params(comp):
    """The seven blocks of an LSTM component as one flat float32 vector in GetParams order (LPS.h:162-189): what
    LstmProjectedStreams.SetParams / lstmp_b200_set_flat take."""
    assert comp.type in LSTM_TYPES
    return np.concatenate([np.asarray(a, np.float32).ravel() for a in comp.arrays])


def lstm_component_from_flat(flat, output_dim, input_dim, ncell, num_stream=None):
    flat = np.asarray(flat, np.float32).ravel()
    need = sum(int(np.prod(shape)) for _, shape in lstm_shapes(output_dim, input_dim, ncell))
    if need != flat.size:
        raise RuntimeError("flat parameter vector has %d elements, the component needs %d" % (flat.size, need))
    arrays, off = [], 0
    for _, shape in lstm_shapes(output_dim, input_dim, ncell):
        n = int(np.prod(shape))
        arrays.append(flat[off:off + n].reshape(shape).copy())
        off += n
    attrs = [("<CellDim>", int(ncell))]
    typ = "<LstmProjected>"
    if num_stream is not None:
        attrs.append(("<NumStream>", int(num_stream)))
        typ = "<LstmProjectedStreams>"
    return NnetComponent(typ, output_dim, input_dim, attrs, arrays)


def time_shift_rows(num_frames, shift):
    """Row gather of TimeShift::PropagateFnc (standard/nnet/nnet-time-shift.h:42-51): out row dst = in row
    clamp(dst + shift, 0, num_frames - 1)."""
    return np.clip(np.arange(num_frames) + int(shift), 0, max(num_frames - 1, 0)).astype(np.int64)
