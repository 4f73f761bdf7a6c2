// TEST INFRASTRUCTURE: stands in for upstream Kaldi base/kaldi-common.h (not vendored by the reference); see ../kaldi-ref-shim.h
#include "kaldi-ref-shim.h"
