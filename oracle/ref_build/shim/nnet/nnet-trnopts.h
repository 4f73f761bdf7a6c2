// TEST INFRASTRUCTURE: stands in for upstream Kaldi nnet/nnet-trnopts.h (not vendored by the reference); see ../kaldi-ref-shim.h
#include "kaldi-ref-shim.h"
