"""B200-native LstmProjectedStreams engine -- Python host-side mirror of the reference component.

Only the hot path of dophist/kaldi-lstm is here: the C-ABI shared library
``_lib/liblstmp_b200.so`` (hand-written sm_100a CUDA, built by ``__graft_entry__.build()``), a
ctypes binding of it (``engine``), a mirror of the reference's component interface
(``component.LstmProjectedStreams``: PropagateFnc / BackpropagateFnc / Update / Reset ...,
google/nnet/bd-nnet-lstm-projected-streams.h) and of the trainer's multi-stream chunk filling
(``dispatch.StreamDispatcher``, google/nnetbin/bd-nnet-train-lstm-streams.cc:128-209).

There is no CPU fallback: importing works anywhere, but creating an engine without the built
library or without a B200 raises.
"""
from .engine import Engine, EngineError, XentEngine, lib_path, load_library  # noqa: F401
from .component import LstmProjectedStreams, NnetTrainOptions, TimeShift  # noqa: F401
from .dispatch import DeviceStreamDispatcher, StreamDispatcher  # noqa: F401
from .loss import Xent, posterior_to_csr  # noqa: F401
from .tail import AffineSoftmaxXent, TailEngine  # noqa: F401
from . import nnet_io, parallel  # noqa: F401
