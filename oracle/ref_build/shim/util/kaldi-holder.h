// TEST INFRASTRUCTURE: stands in for upstream Kaldi util/kaldi-holder.h (not vendored by the reference); see ../kaldi-ref-shim.h
#include "kaldi-ref-shim.h"
