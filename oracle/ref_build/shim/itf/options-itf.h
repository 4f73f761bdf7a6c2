// TEST INFRASTRUCTURE: stands in for upstream Kaldi itf/options-itf.h (not vendored by the reference); see ../kaldi-ref-shim.h
#include "kaldi-ref-shim.h"
