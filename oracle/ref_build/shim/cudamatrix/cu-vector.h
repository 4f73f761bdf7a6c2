// TEST INFRASTRUCTURE: stands in for upstream Kaldi cudamatrix/cu-vector.h (not vendored by the reference); see ../kaldi-ref-shim.h
#include "kaldi-ref-shim.h"
